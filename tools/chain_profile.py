"""Phase profile of the MLP-chain launches of one eager GlobalSPFN step (CPFN_CHAIN_PROFILE=1; the kernels' own
clock64 sums, see csrc/mlp_chain.cu): where the MMA thread, the weight producer and a worker warp of every CTA spend
their time.  Prints one row per launch, cycles averaged per CTA."""
import ctypes
import os
import sys

os.environ["CPFN_CHAIN_PROFILE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from cpfn_b200 import _lib, api, synth  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    eng = api.GlobalSPFN(output_sizes=[3, 4, 28], device=dev)
    sd = {k: torch.from_numpy(v) for k, v in synth.network_state(eng.model.state_dict(), seed=1234).items()}
    eng.load_state_dict(sd, strict=True)
    P = torch.from_numpy(synth.shape_batch(16, 8192, seed=1234, k_slots=28)[0]).to(dev)
    L = _lib.lib()
    for _ in range(3):
        eng.forward(P, dropout=True)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (64 * 16))()
    L.cpfn_debug_chain_profile(buf, 64, 1)
    reps = 5
    for _ in range(reps):
        eng.forward(P, dropout=True)
    n = L.cpfn_debug_chain_profile(buf, 64, 0)
    a = np.array(buf[:], dtype=np.float64).reshape(64, 16)[:n]
    per = n // reps
    a = a.reshape(reps, per, 16).mean(axis=0)
    names = ["SA1", "SA2", "SA3.0", "SA3.1", "SA3.2", "FP1.0", "FP1.1", "FP2", "heads"]
    print("launch            CTAs tiles/CTA | MMA thread: wait-operands wait-weights total | producer: wait-stage total | "
          "worker: load-tile wait-acc total   (kilo-cycles per CTA)")
    for i in range(per):
        r = a[i]
        ct = max(r[8], 1.0)
        k = lambda v: v / ct / 1e3
        print("%-2d %-12s %6.0f %8.2f | %10.1f %10.1f %8.1f | %8.1f %8.1f | %8.1f %8.1f %8.1f | epilogue per layer: %s" % (
            i, names[i] if per == len(names) else "", r[8], r[9] / ct, k(r[0]), k(r[1]), k(r[2]), k(r[3]), k(r[4]),
            k(r[5]), k(r[6]), k(r[7]), " ".join("%.1f" % k(v) for v in r[10:16] if v > 0)))


if __name__ == "__main__":
    main()
