"""The exhaustive proofs behind the kernels' shorter arithmetic forms are programs in tools/:
this runs the quick ones in full -- the YUV->RGB quantiser fusion over every float in [0,1] for all four channel maxima, the
two-operation division over all 2^32 numerators for three divisors including the one it must refuse -- and checks that the
exception list compiled into the library is exactly what the program finds.  (The 16-pair `quant` proof and the other 53 divisors
take minutes: run `tools/verify_quant.c` by hand; the device re-runs the division proof in tests/test_gpu_division.py.)"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "vq")
    subprocess.run(["gcc", "-O2", "-mfma", "-fopenmp", "-ffp-contract=off", os.path.join(ROOT, "tools", "verify_quant.c"), "-o", exe, "-lm"], check=True)
    return exe


def test_rgb_quantiser_fusion_is_exact_on_the_whole_unit_interval(tmp_path):
    out = subprocess.run([_build(tmp_path), "rgbquant"], check=True, capture_output=True, text=True, timeout=600).stdout
    lines = [l for l in out.splitlines() if l.startswith("rgb depth")]
    assert len(lines) == 4 and all(l.endswith("bad=0") for l in lines), out


def test_two_operation_division_and_its_exception(tmp_path):
    hdr = open(os.path.join(ROOT, "libavif_b200", "csrc", "verified_divisors.h")).read()
    refused = re.search(r"kAbNotTwoOpDivisors\[\] = \{([^}]*)\}", hdr).group(1)
    refused = [int(x.strip().rstrip("u"), 16) for x in refused.split(",") if x.strip()]
    assert refused == [0x3FEE5BB7]
    out = subprocess.run([_build(tmp_path), "div", "3fc9930c", "3fee5bb7", "477fff00"], check=True, capture_output=True, text=True, timeout=600).stdout
    bad = {int(l.split()[0], 16): int(l.split("bad=")[1]) for l in out.splitlines() if "bad=" in l}
    assert bad[0x3FC9930C] == 0 and bad[0x477FFF00] == 0
    assert bad[0x3FEE5BB7] > 0  # the listed exception really fails: it keeps the three-operation form


def test_level_order_threshold_search_equals_a_sorted_search():
    """csrc/tonemap.cu keeps the 2^d - 1 output-transfer-function thresholds of a depth in LEVEL order (Eytzinger) in shared memory and
    walks i -> 2i + (E[i] <= x); the leaf reached minus 2^d must be the number of thresholds <= x (what a search of the sorted
    array gives), NaN thresholds (codes the function never reaches) acting as +inf.  Restated in numpy for d = 8, 10, 12."""
    import numpy as np
    rng = np.random.RandomState(3)
    for d in (8, 10, 12):
        n = (1 << d) - 1
        t = np.sort(rng.random_sample(n).astype(np.float32))
        t[-5:] = np.nan  # unreachable codes sit at the tail
        i = np.arange(1, n + 1)
        k = np.floor(np.log2(i)).astype(np.int64)
        j = i - (1 << k)
        src = ((2 * j + 1) << (d - 1 - k)) - 1  # eytzinger_source()
        assert np.array_equal(np.sort(src), np.arange(n))
        e = np.empty(n + 1, dtype=np.float32)
        e[0] = 0
        e[1:] = t[src]
        x = np.concatenate([rng.random_sample(5000).astype(np.float32) * 1.2 - 0.1, t[:200], [np.float32(np.inf), np.float32(-np.inf)]])
        pos = np.ones(len(x), dtype=np.int64)
        with np.errstate(invalid="ignore"):
            for _ in range(d):
                pos = 2 * pos + (e[pos] <= x)
        want = np.searchsorted(t[:-5], x, side="right")  # NaN tail never counts
        assert np.array_equal(pos - (1 << d), want), d


def test_output_code_is_a_step_function_of_the_linear_value_except_where_listed():
    """Premise of the tone mapper's threshold table (csrc/tonemap.cu): for integer outputs the stored code
    (T)(0.5 + clamp(linearToGamma(x)) * max) never decreases as x grows and is 0 for x <= 0.  Scanned on the compiled reference's own
    thirteen functions (avifTransferCharacteristicsGetLinearToGammaFunction, src/colr.c:216-491) over 1.2 M floats from -1e5 to 1e5:
    it holds for every function except BT.1361 extended (its negative branch is positive just below zero), and the list of exclusions
    compiled into the library must be exactly that set."""
    import ctypes
    import numpy as np
    from oracle import loader
    ref = loader.load_reference()
    raw = getattr(ref, "raw", ref)
    if raw is None or not hasattr(raw, "avifTransferCharacteristicsGetLinearToGammaFunction"):
        import pytest
        pytest.skip("compiled reference (oracle/_ref) not available")
    src = open(os.path.join(ROOT, "libavif_b200", "csrc", "tonemap.cu")).read()
    listed = sorted(int(x) for x in re.search(r"kAbNotStepFunctions\[\] = \{([^}]*)\}", src).group(1).split(",") if x.strip())
    fn_t = ctypes.CFUNCTYPE(ctypes.c_float, ctypes.c_float)
    get = raw.avifTransferCharacteristicsGetLinearToGammaFunction
    get.restype, get.argtypes = fn_t, [ctypes.c_int]
    pos = np.geomspace(1e-12, 1e5, 400000)
    xs = np.unique(np.concatenate([-pos, [0.0], pos, np.linspace(0, 1.2, 200001)]).astype(np.float32))
    failing = []
    for tc in [1, 4, 5, 7, 8, 9, 10, 11, 12, 13, 16, 17, 18]:
        f = get(tc)
        g = np.array([f(float(x)) for x in xs], dtype=np.float32)
        g = np.minimum(np.float32(1.0), np.maximum(np.float32(0.0), g))
        ok = True
        for mx in (255.0, 1023.0, 4095.0):
            code = np.floor(np.float32(0.5) + g * np.float32(mx)).astype(np.int64)
            ok = ok and not (np.diff(code) < 0).any() and int(code[xs <= 0].max()) == 0
        if not ok:
            failing.append(tc)
    assert failing == listed == [12], (failing, listed)


def test_threshold_search_reproduces_the_reference_code_function():
    """The tone mapper's table algorithm restated on the CPU with the compiled reference's own float functions: thresholds by bisection
    over positive float bit patterns (smallest x whose code is >= k), code(x) = number of thresholds <= x.  For every transfer function
    the library serves this way the search must give EXACTLY the code the reference's formula gives, on random, log-spaced and negative
    inputs (the device evaluates the function in double where the reference uses float libm: that difference, not the algorithm, is
    what the +-1-code tolerance of tests/test_gpu_gainmap.py covers)."""
    import ctypes
    import numpy as np
    from oracle import loader
    raw = getattr(loader.load_reference(), "raw", None)
    if raw is None or not hasattr(raw, "avifTransferCharacteristicsGetLinearToGammaFunction"):
        import pytest
        pytest.skip("compiled reference (oracle/_ref) not available")
    fn_t = ctypes.CFUNCTYPE(ctypes.c_float, ctypes.c_float)
    get = raw.avifTransferCharacteristicsGetLinearToGammaFunction
    get.restype, get.argtypes = fn_t, [ctypes.c_int]
    rng = np.random.RandomState(1)

    def as_float(b):
        return np.array([b], dtype=np.uint32).view(np.float32)[0]
    for tc, depth in [(1, 10), (4, 10), (5, 8), (7, 10), (8, 10), (9, 10), (10, 8), (11, 10), (13, 12), (13, 8), (16, 10), (17, 10), (18, 10)]:
        f, mx = get(tc), np.float32((1 << depth) - 1)

        def code(x):
            g = min(np.float32(1.0), max(np.float32(0.0), np.float32(f(float(x)))))
            return int(np.floor(np.float32(0.5) + g * mx))
        top = code(np.float32(np.inf))
        thresholds = []
        for k in range(1, int(mx) + 1):
            if top < k:
                break  # codes the function never reaches: nothing is ever >= their thresholds
            lo, hi = 0, 0x7F800000
            while hi - lo > 1:
                mid = lo + ((hi - lo) >> 1)
                if code(as_float(mid)) >= k:
                    hi = mid
                else:
                    lo = mid
            thresholds.append(as_float(hi))
        t = np.array(thresholds, dtype=np.float32)
        assert np.all(np.diff(t) >= 0)
        xs = np.concatenate([rng.random_sample(6000).astype(np.float32) * 1.5, (10.0 ** rng.uniform(-8, 2, 6000)).astype(np.float32),
                             -rng.random_sample(500).astype(np.float32), t[:: max(1, len(t) // 200)]])
        want = np.array([code(x) for x in xs])
        got = np.searchsorted(t, xs, side="right")
        assert np.array_equal(want, got), (tc, depth, int(np.abs(want - got).max()))
