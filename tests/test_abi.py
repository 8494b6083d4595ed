"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol the
public header declares, and refuses to compute without a CUDA device (no fallback)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "gcge_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_something():
    syms = declared_symbols()
    assert len(syms) >= 25, syms


def test_library_exports_every_declared_symbol():
    from gcge_b200 import api
    L = api.lib()
    missing = [s for s in declared_symbols() if not hasattr(L, s)]
    assert not missing, f"declared in include/gcge_b200.h but not exported: {missing}"


def test_header_cites_reference_interfaces():
    text = (ROOT / "include" / "gcge_b200.h").read_text()
    # every slot-level entry point names the reference file:line it replaces
    for needle in ("app/app_ccs.c:50-139", "app/app_lapack.c:334-395", "app/app_lapack.c:463-534",
                   "src/ops_multi_vec.c:351-411", "src/ops_orth.c:203-393", "src/ops_lin_sol.c:140-437",
                   "src/ops_eig_sol_gcg.c:1253-1558", "src/ops_eig_sol_gcg.c:1201-1204"):
        assert needle in text, needle


def test_no_cpu_fallback_without_device():
    from gcge_b200 import api
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present; the failure path is exercised on the CPU box")
    with pytest.raises(api.B200Error, match="no CUDA device|no CPU fallback"):
        api.init(0)
    with pytest.raises(api.B200Error):
        api.MultiVec(4, 2)


def test_ops_adaptor_exports():
    p = ROOT / "gcge_b200" / "lib" / "libgcge_b200_ops.so"
    if not p.exists():
        pytest.skip("OPS adaptor is built only where the reference headers exist")
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", str(p)], capture_output=True, text=True).stdout
    for sym in ("OPS_B200_Set", "B200_MatCreateFromCCS", "B200_MatDestroy", "EigenSolverSetup_GCG_B200",
                "MultiLinearSolverSetup_BlockPCG_B200", "MultiVecOrthSetup_ModifiedGramSchmidt_B200"):
        assert re.search(rf"\bT {sym}\b", out), sym


def test_product_does_not_touch_oracle():
    """Nothing under gcge_b200/ or include/ may import, link or call the oracle."""
    bad = []
    for p in list((ROOT / "gcge_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if p.suffix in {".py", ".c", ".h", ".cu", ".cuh"} or p.name == "Makefile":
            t = p.read_text(errors="ignore")
            if re.search(r"from\s+oracle|import\s+oracle|oracle/_ref|libgcge_ref|libgcge_oracle|gcg_numpy", t):
                bad.append(str(p))
    assert not bad, bad


@pytest.mark.parametrize("seed,steps", [(0, 0), (0, 1), (1, 30), (7, 31), (0, 1984), (123, 123457), (0, 3000000)])
def test_rand_jump_ahead_matches_glibc(seed, steps):
    """Host half of the device MultiVecSetRandomValue (reference app/app_lapack.c:322-333): the
    31x31 matrix-power jump of glibc's TYPE_3 generator, the state read-out through
    initstate()/setstate() and the write-back -- checked against glibc's own rand()."""
    from gcge_b200 import api
    L = api.lib()
    L.b200_rand_selfcheck.argtypes = [C.c_ulonglong]
    C.CDLL("libc.so.6").srand(C.c_uint(seed))
    assert L.b200_rand_selfcheck(steps) == 0, L.b200_last_error()


def test_rand_exact_jump_with_binary_powers():
    """The device fill (b200_rand.cu) gives every lane its own stream position t = column * n + row and builds the
    generator state there as M^t S_0 from the powers M^(2^j) of the 31 x 31 step matrix over Z/2^32.  The same
    arithmetic restated in numpy against glibc itself: rand() called t times, then the next values compared."""
    import numpy as np
    DEG, SEP = 31, 3
    libc = C.CDLL("libc.so.6")
    M = np.zeros((DEG, DEG), dtype=np.uint64)
    for i in range(DEG - 1):
        M[i, i + 1] = 1
    M[DEG - 1, 0] = 1; M[DEG - 1, DEG - SEP] = 1                 # o_t = o_{t-31} + o_{t-3}
    mask = np.uint64(0xFFFFFFFF)

    def mm(a, b):
        # 32-bit wrap-around products of 31 terms: split b into 16-bit halves so nothing overflows 64 bits
        lo = (a @ (b & np.uint64(0xFFFF))) & mask
        hi = (a @ (b >> np.uint64(16))) & np.uint64(0xFFFF)
        return (lo + (hi << np.uint64(16))) & mask

    powers = [M]
    for _ in range(1, 24):
        powers.append(mm(powers[-1], powers[-1]))
    # S_0: the 31 outputs before position 0, recovered from glibc by running the recurrence backwards is not needed:
    # take the history as the first 31 FULL words o_t, which rand() hides one bit of -- so start from a state we know:
    # after srand(seed) glibc's table is r[i] (i < 34 by the LCG, then 310 discarded steps); rebuild it as glibc does.
    seed = 12345
    r = [0] * 34
    r[0] = seed
    for i in range(1, 31):
        hi_, lo_ = divmod(r[i - 1], 127773)
        w = 16807 * lo_ - 2836 * hi_
        r[i] = w + 2147483647 if w < 0 else w
    for i in range(31, 34):
        r[i] = r[i - 31]
    o = [(x & 0xFFFFFFFF) for x in r]
    for i in range(34, 344):
        o.append((o[i - 31] + o[i - 3]) & 0xFFFFFFFF)
    S0 = np.array(o[344 - 31:344], dtype=np.uint64)               # history right before the first rand() after srand
    for t in (0, 1, 5, 31, 1000, 65537, 3_000_001):
        S = S0.copy()
        for j in range(24):
            if (t >> j) & 1:
                S = mm(powers[j], S.reshape(DEG, 1)).reshape(DEG)
        libc.srand(C.c_uint(seed))
        for _ in range(t):
            libc.rand()
        s = [int(v) for v in S]
        for i in range(40):                                       # the recurrence as the kernel runs it
            nxt = (s[0] + s[DEG - SEP]) & 0xFFFFFFFF
            s = s[1:] + [nxt]
            assert (nxt >> 1) == libc.rand(), (t, i)
