"""CPU oracle for the CPFN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this package.  Nothing
under ``cpfn_b200/`` imports it; the product path fails loudly when its CUDA
library is missing instead of falling back to anything here.

Contents
--------
``cpfn_oracle.c`` / ``index_ops.py``  bit-exact C restatement of the nine
    pointnet2 kernels (FPS, ball query, gather/group (+grad), 3-NN,
    3-weighted-sum (+grad)).
``fitters.py``   numpy fp32 restatement of SPFN's weighted-TLS fitters.
``network.py``   plain torch-fp32 CPU restatement of the SA / FP layers and the
    PointNet2 forward, built on ``index_ops``.
``build_ref.py`` recipe that compiles the UNMODIFIED reference CUDA extension
    from /root/reference into ``oracle/_ref/`` (git-ignored) for live parity
    runs on the GPU box.

Parity pin status is recorded per module in its header and in DESIGN.md.
"""
