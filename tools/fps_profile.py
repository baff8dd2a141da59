import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpfn_b200 import _lib, cuda_ops, synth
dev = torch.device("cuda:0")
P = torch.from_numpy(synth.shape_batch(16, 8192, seed=1234)[0]).to(dev)
os.environ["CPFN_FPS_PROFILE"] = "1"
for C in (4, 8, 2):
    os.environ["CPFN_FPS_CLUSTER"] = str(C)
    for _ in range(2):
        cuda_ops.farthest_point_sampling(P, 512)
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 6)()
    _lib.check(_lib.lib().cpfn_debug_fps_profile(ctypes.cast(buf, ctypes.c_void_p)), "prof")
    v = np.array(list(buf), dtype=np.float64) / 511
    print("C", C, "cycles/round: update %.0f | warp argmax %.0f | push %.0f | wait %.0f | cluster argmax+unrank %.0f | centroid %.0f | total %.0f" % (*v, v.sum()))
