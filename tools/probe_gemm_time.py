"""In-kernel phase timing of the tcgen05 GEMM (clock64 stamps) for the three head passes."""
import ctypes as C, sys
import torch
sys.path.insert(0, ".")
from acezero_b200 import _lib
lib = _lib.load()
NAMES = ["prologue", "dep wait", "first tile", "mainloop issue", "-> acc ready(from first tile)", "epilogue", "exit"]

def run(tag, a_mn, b_mn, M, N, K, batch, epi, bn):
    g = torch.Generator(device="cuda").manual_seed(1)
    A = (torch.randn((batch, K, M) if a_mn else (batch, M, K), device="cuda", generator=g) * 0.5).half()
    B = (torch.randn((batch, K, N) if b_mn else (batch, N, K), device="cuda", generator=g) * 0.5).half()
    d = _lib.GemmDesc()
    d.A, d.B, d.a_mn_major, d.b_mn_major = A.data_ptr(), B.data_ptr(), a_mn, b_mn
    d.M, d.N, d.K, d.batch, d.bn, d.epilogue = M, N, K, batch, bn, epi
    d.a_zstride, d.b_zstride, d.lda, d.ldb = A.stride(0), B.stride(0), A.stride(1), B.stride(1)
    if epi == 2:
        out = torch.empty((batch, M, N), device="cuda"); d.out32, d.out32_zstride, d.ldo32 = out.data_ptr(), M * N, N
    else:
        out = torch.empty((M, N), device="cuda", dtype=torch.float16); d.out, d.ldo, d.relu = out.data_ptr(), N, 1
        bias = torch.zeros(N, device="cuda"); d.bias = bias.data_ptr()
        if epi == 1:
            mask = torch.ones((M, N), device="cuda", dtype=torch.float16); d.mask = mask.data_ptr()
    ctas = ((N + bn - 1) // bn) * ((M + 127) // 128) * batch
    clk = torch.zeros((ctas, 8), device="cuda", dtype=torch.int64)
    for it in range(6):
        d.dbg_clock = clk.data_ptr() if it == 5 else None
        _lib.check(lib.acez_gemm_f16(C.byref(d), _lib.stream_ptr()))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    d.dbg_clock = None
    e0.record()
    for _ in range(20): _lib.check(lib.acez_gemm_f16(C.byref(d), _lib.stream_ptr()))
    e1.record(); torch.cuda.synchronize()
    c = clk.double().cpu()
    t = c - c[:, :1]
    m = t.mean(0)
    print(f"{tag}: {ctas} CTAs, {e0.elapsed_time(e1) / 20 * 1000:.2f} us/launch (back-to-back incl. host encode)")
    print("   mean cycles since CTA entry: prologue %.0f | dep %.0f | first tile %.0f | mma issued %.0f | acc ready %.0f | epi done %.0f | exit %.0f"
          % (m[1], m[2], m[3], m[4], m[5], m[6], m[7]))
    print("   max exit %.0f cycles = %.2f us @1.965GHz" % (t[:, 7].max(), t[:, 7].max() / 1965))

run("fwd   5120x512x512 BN256", 0, 0, 5120, 512, 512, 1, 0, 256)
run("fwd   5120x512x512 BN128", 0, 0, 5120, 512, 512, 1, 0, 128)
run("dgrad 5120x512x512 BN256", 0, 1, 5120, 512, 512, 1, 1, 256)
run("wgrad 8x512x512x5120 BN128", 1, 1, 512, 512, 5120, 8, 2, 128)
