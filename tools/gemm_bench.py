"""Times the DMMA GEMM engine on the shapes the DGP path uses (through dgp_debug_gemm). A/B: DGP_B200_LIB=<other .so>."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dgp_toolbox_b200 as D

ctx = D._lib.get_context(0)
P = 131072
cases = [  # name, nt, M, N, K, tri, clower, batch, splitk
    ("V=Linv K      (NN tri lower)", 0, 256, P, 256, 1, 0, 1, 1),
    ("A=LinvT V     (NN tri upper)", 0, 256, P, 256, 2, 0, 1, 1),
    ("T=RpT A  x8   (NN tri upper batched)", 0, 256, P, 256, 2, 0, 8, 1),
    ("W=Kinv dA     (NN dense)", 0, 256, P, 256, 0, 0, 1, 1),
    ("dKu=-Wg A^T   (NT, split-K)", 1, 256, 256, P, 0, 0, 1, 27),
    ("dR=A s T^T x8 (NT c_lower, kscale, split-K)", 1, 256, 256, P, 0, 1, 8, 11),
    ("dR no kscale  (NT c_lower, split-K)", 1, 256, 256, P, 0, 1, 8, 11),
    ("dR splitk 14", 1, 256, 256, P, 0, 1, 8, 14),
    ("dR splitk 18", 1, 256, 256, P, 0, 1, 8, 18),
    ("dR splitk 22", 1, 256, 256, P, 0, 1, 8, 22),
    ("dR splitk 33", 1, 256, 256, P, 0, 1, 8, 33),
    ("dR splitk 44", 1, 256, 256, P, 0, 1, 8, 44),
    ("dqmu=A Gm     (NN N=32 split-K)", 0, 256, 32, P, 0, 0, 1, 32),
]
for name, nt, M, N, K, tri, clow, batch, sk in cases:
    A = torch.randn(batch, M, K, dtype=torch.float64, device="cuda")
    B = torch.randn(batch, N, K, dtype=torch.float64, device="cuda") if nt else torch.randn(batch, K, N, dtype=torch.float64, device="cuda")
    C = torch.zeros(batch, M, N, dtype=torch.float64, device="cuda")
    ks = torch.randn(batch, K, dtype=torch.float64, device="cuda") if (nt and clow and "x8" in name) else None
    def run():
        ctx.call("dgp_debug_gemm", nt, M, N, K, 1.0, D._lib.ptr(A), D._lib.ptr(B), 0.0, D._lib.ptr(C), tri, clow, batch, sk, D._lib.ptr(ks))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    dense = 2.0 * M * N * K * batch
    useful = dense * (0.5 if (tri or clow) else 1.0)
    print(f"{name:46s} {ms:8.3f} ms  dense-equivalent {dense/ms/1e9:6.2f} TF  useful {useful/ms/1e9:6.2f} TF")
