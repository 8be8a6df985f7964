"""GPU bring-up harness for the tcgen05 GEMM: numerics vs fp32 torch, timing vs cuBLAS bf16.

Each case runs in its own subprocess under a timeout so that a hung kernel cannot take the whole gpurun call down.
Usage:  python scripts/gemm_check.py            (driver: runs all cases, writes gpurun_out/gemm_check.json)
        python scripts/gemm_check.py --case N   (one case, prints a JSON line)
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

# (name, M, N, K, a_mn, b_mn, out_f32, epi, splits)
CASES = [
    ("fwd_small", 128, 256, 64, 0, 0, 0, 0, 1),
    ("fwd_k256", 128, 256, 256, 0, 0, 0, 0, 1),
    ("fwd_multi", 512, 768, 768, 0, 0, 0, 0, 1),
    ("fwd_ragged", 300, 1000, 200, 0, 0, 0, 0, 1),
    ("fwd_bias", 512, 768, 768, 0, 0, 0, 1, 1),
    ("fwd_bias_gelu", 512, 3072, 768, 0, 0, 0, 2, 1),
    ("fwd_bias_resid", 512, 768, 3072, 0, 0, 0, 3, 1),
    ("dgrad_bmn", 512, 768, 3072, 0, 1, 0, 0, 1),
    ("dgrad_dgelu", 512, 3072, 768, 0, 1, 0, 4, 1),
    ("wgrad_mnmn_bf16", 768, 768, 512, 1, 1, 0, 0, 1),
    ("wgrad_f32", 768, 768, 2048, 1, 1, 1, 0, 1),
    ("wgrad_f32_split", 768, 3072, 4096, 1, 1, 1, 0, 8),
    ("amn_bk", 256, 512, 512, 1, 0, 0, 0, 1),
    ("perf_fwd_qkv", 16384, 2304, 768, 0, 0, 0, 1, 1),
    ("perf_fwd_fc", 16384, 3072, 768, 0, 0, 0, 2, 1),
    ("perf_fwd_proj", 16384, 768, 3072, 0, 0, 0, 3, 1),
    ("perf_dgrad_fc", 16384, 768, 3072, 0, 1, 0, 0, 1),
    ("perf_wgrad_fc", 3072, 768, 16384, 1, 1, 1, 0, 8),
    ("perf_lmhead", 8192, 50258, 768, 0, 0, 0, 0, 1),
    ("perf_square", 8192, 8192, 8192, 0, 0, 0, 0, 1),
]


def run_case(idx: int) -> dict:
    import torch

    from distributedtraining_b200.ops import _lib

    name, M, N, K, a_mn, b_mn, out_f32, epi, splits = CASES[idx]
    torch.manual_seed(idx)
    dev = "cuda"
    L = _lib.lib()
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()  # logical A[M,K]
    B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()  # logical B[N,K]; C = A @ B^T
    A_st = A.t().contiguous() if a_mn else A  # MN-major storage: [K, M]
    B_st = B.t().contiguous() if b_mn else B
    lda = A_st.shape[1]
    ldb = B_st.shape[1]
    ldc = (N + 7) // 8 * 8
    bias = (torch.randn(N, device=dev) * 0.5).bfloat16()
    aux = (torch.randn(M, ldc, device=dev) * 0.5).bfloat16()
    C = torch.zeros(M, ldc, device=dev, dtype=torch.float32 if out_f32 else torch.bfloat16)
    C2 = torch.zeros(M, ldc, device=dev, dtype=torch.bfloat16) if epi == 2 else None

    def call():
        rc = L.dtb_gemm_bf16(
            _lib.ptr(A_st), _lib.ptr(B_st), _lib.ptr(C), M, N, K, lda, ldb, ldc, a_mn, b_mn, out_f32, epi,
            _lib.ptr(bias), _lib.ptr(aux), ldc, _lib.ptr(C2), ldc, ctypes.c_float(1.0), splits, _lib.num_sms(),
            _lib.stream_ptr(), None, 0, None, 0, None, 0, ctypes.c_float(0.0), None)
        assert rc == 0, rc

    call()
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    if epi in (1, 2, 3):
        ref = ref + bias.float()[None, :]
    pre = ref
    if epi == 2:
        ref = torch.nn.functional.gelu(pre, approximate="tanh")
    if epi in (3, 5):
        ref = ref + aux[:, :N].float()
    if epi == 4:
        x = aux[:, :N].float().requires_grad_(True)
        g = torch.autograd.grad(torch.nn.functional.gelu(x, approximate="tanh").sum(), x)[0]
        ref = ref * g
    got = C[:, :N].float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    out = {"name": name, "M": M, "N": N, "K": K, "max_abs_err": err, "ref_max": scale, "rel": err / max(scale, 1e-6)}
    if epi == 2:
        out["rel_pre"] = ((C2[:, :N].float() - pre).abs().max() / pre.abs().max()).item()
    # timing
    if M * N * K >= 2 ** 31:
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
        for _ in range(3):
            if out_f32:
                C.zero_()
            call()
        ts = []
        for _ in range(10):
            flush.zero_()
            if out_f32:
                C.zero_()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            call()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        out["ms"] = ms
        out["tflops"] = 2.0 * M * N * K / ms / 1e9
        tt = []
        Bt = B.t()
        for _ in range(3):
            torch.matmul(A, Bt)
        for _ in range(10):
            flush.zero_()
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(A, Bt)
            e1.record()
            torch.cuda.synchronize()
            tt.append(e0.elapsed_time(e1))
        out["cublas_ms"] = sorted(tt)[len(tt) // 2]
        out["cublas_tflops"] = 2.0 * M * N * K / out["cublas_ms"] / 1e9
    out["ok"] = bool(out["rel"] < 2e-2)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", type=int, default=-1)
    ap.add_argument("--only", type=str, default="")
    args = ap.parse_args()
    if args.case >= 0:
        print("RESULT " + json.dumps(run_case(args.case)), flush=True)
        return
    os.makedirs("gpurun_out", exist_ok=True)
    results = []
    for i, c in enumerate(CASES):
        if args.only and args.only not in c[0]:
            continue
        try:
            r = subprocess.run([sys.executable, __file__, "--case", str(i)], capture_output=True, text=True, timeout=120)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            if line:
                res = json.loads(line[-1][7:])
            else:
                res = {"name": c[0], "ok": False, "error": (r.stderr or r.stdout)[-600:]}
        except subprocess.TimeoutExpired:
            res = {"name": c[0], "ok": False, "error": "TIMEOUT (hang)"}
        print(json.dumps(res), flush=True)
        results.append(res)
        with open("gpurun_out/gemm_check.json", "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
