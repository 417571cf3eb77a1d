"""The C-ABI library loads on a CPU-only box, exports every symbol include/aus_b200.h declares, and its pure-host
helpers (frame count, windows, rfftfreq, shard plan) agree with the oracle.  No compute call is made here."""
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SR = 44100


def _declared():
    h = open(os.path.join(ROOT, "include", "aus_b200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(aus_[a-z0-9_]+)\s*\(", h)))


def test_every_declared_symbol_is_exported_and_bound(aus):
    from aus_rust_b200 import _lib
    lib = aus.load_library()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/aus_b200.h but not exported by libaus_b200.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert sorted(_lib.SIGNATURES) == names
    assert b"sm_100a" in lib.aus_version()


def test_library_contains_only_sm100a_code():
    import subprocess
    so = os.path.join(ROOT, "aus-rust_b200", "csrc", "libaus_b200.so")
    out = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("N,hop", [(64, 16), (1024, 256), (2048, 512), (2048, 2048), (16384, 4096), (2048, 1)])
def test_num_frames_matches_oracle(aus, orc, N, hop):
    rng = np.random.default_rng(N + hop)
    for n in list(range(0, 40)) + [N - 1, N, N + 1, N + hop, N + hop + 1] + list(rng.integers(0, 50 * N, 200)):
        assert aus.num_frames(int(n), N, hop) == orc.num_frames(int(n), N, hop)
    assert aus.irstft_len(7, N, hop) == N + 6 * hop
    assert aus.batch_num_frames([0, N, 3 * N + 1], N, hop) == sum(orc.num_frames(n, N, hop) for n in (0, N, 3 * N + 1))


@pytest.mark.parametrize("wt", range(5))
def test_windows_match_oracle(aus, orc, wt):
    for L in (1, 2, 5, 16, 17, 1024, 2032):
        got, ref = aus.generate_window(wt, L), orc.generate_window(wt, L)
        assert np.array_equal(np.isnan(got), np.isnan(ref))
        assert np.allclose(got, ref, rtol=0, atol=2e-16, equal_nan=True)     # libm cos vs numpy cos: <= 1 ulp
    assert [w.name for w in aus.WindowType] == ["Bartlett", "Blackman", "Hanning", "Hamming", "Rectangular"]   # window.rs:9-15


def test_rfftfreq_matches_oracle(aus, orc):
    for N, sr in ((2048, 44100), (1024, 48000), (16384, 22050)):
        assert np.array_equal(aus.spectrum.rfftfreq(N, sr), orc.rfftfreq(N, sr))


def test_mirror_has_the_reference_api(aus):
    """Same names and argument order as the crate's public functions on this path."""
    want = {
        aus.spectrum: {"rfft": ["audio", "fft_size"], "irfft": ["spectrum", "fft_size"],
                       "rstft": ["audio", "fft_size", "hop_size", "window_type"],
                       "irstft": ["spectrogram", "fft_size", "hop_size", "window_type"],
                       "complex_to_polar_rfft": ["spectrum"], "complex_to_polar_rstft": ["spectrogram"],
                       "polar_to_complex_rfft": ["magnitude_spectrum", "phase_spectrum"],
                       "polar_to_complex_rstft": ["magnitude_spectrogram", "phase_spectrogram"],
                       "rfftfreq": ["fft_size", "sample_rate"],
                       "overlap_add": ["audio_chunks", "fft_size", "hop_size"]},
        aus.analysis: {"analyzer": ["magnitude_spectrum", "fft_size", "sample_rate"],
                       "spectral_centroid": ["magnitude_spectrum", "rfft_freqs"], "spectral_entropy": ["magnitude_spectrum"],
                       "spectral_flatness": ["magnitude_spectrum"], "spectral_kurtosis": ["magnitude_spectrum", "rfft_freqs"],
                       "spectral_roll_off_point": ["magnitude_spectrum", "rfft_freqs", "n"],
                       "spectral_skewness": ["magnitude_spectrum", "rfft_freqs"], "spectral_slope": ["magnitude_spectrum"],
                       "spectral_slope_region": ["magnitude_spectrum", "rfft_freqs", "f_lower", "f_upper", "sample_rate"],
                       "spectral_variance": ["magnitude_spectrum", "rfft_freqs"], "make_power_spectrum": ["magnitude_spectrum"],
                       # the public building blocks (spectral_analysis_tools.rs:79-290), argument names as upstream
                       "make_spectrum_pmf": ["power_spectrum", "power_spectrum_sum"], "dot_product": ["vec1", "vec2"],
                       "compute_spectral_centroid": ["magnitude_spectrum", "rfft_freqs", "magnitude_spectrum_sum"],
                       "compute_spectral_entropy": ["spectrum_pmf"],
                       "compute_spectral_flatness": ["magnitude_spectrum", "magnitude_spectrum_sum"],
                       "compute_spectral_kurtosis": ["spectrum_pmf", "rfft_freqs", "spectral_centroid", "spectral_variance"],
                       "compute_spectral_roll_off_point": ["power_spectrum", "rfft_freqs", "power_spectrum_sum", "n"],
                       "compute_spectral_skewness": ["spectrum_pmf", "rfft_freqs", "spectral_centroid", "spectral_variance"],
                       "compute_spectral_slope": ["power_spectrum", "power_spectrum_sum"],
                       "compute_spectral_slope_region": ["power_spectrum", "rfft_freqs", "f_lower", "f_upper", "sample_rate"],
                       "compute_spectral_variance": ["spectrum_pmf", "rfft_freqs", "spectral_centroid"],
                       "autocorrelation": ["audio", "fft_size"]},
        aus.mp: {"stft_analysis": ["audio", "fft_size", "sample_rate", "max_num_threads"]},
    }
    for mod, fns in want.items():
        for name, args in fns.items():
            assert list(inspect.signature(getattr(mod, name)).parameters)[:len(args)] == args, (mod.__name__, name)
    from dataclasses import fields
    assert [f.name for f in fields(aus.analysis.Analysis)] == list(aus.FEATURE_NAMES)      # analyzer.rs:10-23
    for n in ("generate_window_bartlett", "generate_window_blackman", "generate_window_hanning",
              "generate_window_hamming", "generate_window_rectangular", "generate_window"):
        assert callable(getattr(aus, n))


def test_no_cpu_fallback(aus):
    """Without a CUDA device the product path must fail loudly, never compute on the host."""
    lib = aus.load_library()
    if lib.aus_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(aus.AusError) as e:
        aus.Context(0)
    assert e.value.code == -2
    src = open(os.path.join(ROOT, "aus-rust_b200", "api.py")).read() + open(os.path.join(ROOT, "aus-rust_b200", "_lib.py")).read()
    assert "oracle" not in src.replace("aus_oracle", "oracle") or "import aus_oracle" not in src
    for f in os.listdir(os.path.join(ROOT, "aus-rust_b200")):
        if f.endswith(".py"):
            assert "aus_oracle" not in open(os.path.join(ROOT, "aus-rust_b200", f)).read(), f"{f} references the oracle"


def test_only_tests_smoke_and_bench_touch_the_oracle():
    """oracle/ is test infrastructure: nothing under the package or tools/ may import, link or execute it."""
    import re
    pat = re.compile(r"(import\s+aus_oracle|from\s+aus_oracle|libaus_oracle|oracle/_ref|sys\.path[^\n]*oracle)")
    for top in ("aus-rust_b200", "tools"):
        for d, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".rs", "Makefile")):
                    src = open(os.path.join(d, f), errors="replace").read()
                    assert not pat.search(src), f"{os.path.join(d, f)} reaches into oracle/"


# ---------------------------------------------------------------------------------------------- shard plan (pure host)
@pytest.mark.parametrize("n_dev", [1, 2, 4, 8])
def test_shard_plan_single_long_file_reproduces_whole_file_features(aus, orc, n_dev):
    """Cutting one file into contiguous frame ranges with an fft_size-hop halo, and running each range as its own
    file, gives exactly the whole-file result -- including the short last frame, which only the final range sees."""
    N, hop = 512, 128
    x = orc.synth_signal(40 * hop + 333, seed=12)
    ref = orc.stft_features(x, N, hop, orc.HANNING, SR)
    vo, vl, vr, vd = aus.plan_shards([0], [x.size], N, hop, n_dev)
    assert len(vo) == min(n_dev, ref.shape[0]) and sorted(set(vd)) == list(range(len(vo)))
    got = np.full_like(ref, np.nan)
    for o, l, r in zip(vo, vl, vr):
        part = orc.stft_features(x[o:o + l], N, hop, orc.HANNING, SR)
        got[r:r + part.shape[0]] = part
    assert np.array_equal(got, ref)                    # bit-identical: same samples, same windows, same order
    assert np.all(np.diff(vd) >= 0)


def test_shard_plan_many_files(aus, orc):
    rng = np.random.default_rng(3)
    lens = rng.integers(0, 20000, 37)
    offs = np.concatenate([[0], np.cumsum(lens + rng.integers(0, 5, 37))[:-1]])
    for n_dev in (2, 3, 8):
        vo, vl, vr, vd = aus.plan_shards(offs, lens, 1024, 256, n_dev)
        assert np.array_equal(vo, offs) and np.array_equal(vl, lens)          # enough files: whole files per device
        frames = np.array([orc.num_frames(int(n), 1024, 256) for n in lens])
        assert np.array_equal(vr, np.concatenate([[0], np.cumsum(frames)[:-1]]))
        assert np.all(np.diff(vd) >= 0) and vd.min() == 0 and vd.max() == n_dev - 1
        per_dev = np.array([frames[vd == d].sum() for d in range(n_dev)])
        assert per_dev.max() <= frames.sum() / n_dev + frames.max()           # balanced to within one file


def test_shard_plan_properties(aus, orc):
    """Property test of the pure-host planner (hypothesis): for any batch of files, fft_size, hop and device count the
    virtual files (a) stay inside their source file, (b) tile the frame rows 0..total without gap or overlap, (c) each
    produce exactly the rows they claim under the reference's frame-count rule, and (d) go to devices in ascending
    contiguous runs."""
    hyp = pytest.importorskip("hypothesis")
    st = pytest.importorskip("hypothesis.strategies")

    @hyp.settings(max_examples=150, deadline=None)
    @hyp.given(st.lists(st.integers(0, 60000), min_size=1, max_size=12), st.sampled_from([64, 256, 1024, 4096]),
               st.integers(1, 4096), st.integers(1, 9), st.integers(0, 7))
    def run(lens, N, hop, n_dev, gap):
        hop = min(hop, N)
        lens = np.asarray(lens, dtype=np.uint64)
        offs = np.concatenate([[0], np.cumsum(lens + np.uint64(gap))[:-1]]).astype(np.uint64)
        vo, vl, vr, vd = aus.plan_shards(offs, lens, N, hop, n_dev)
        frames = [orc.num_frames(int(n), N, hop) for n in lens]
        starts = np.concatenate([[0], np.cumsum(frames)])
        assert len(vo) >= len(lens) and np.all(np.diff(vd) >= 0) and vd.min() >= 0 and vd.max() < n_dev
        row = 0
        for o, l, r in zip(vo, vl, vr):
            f = int(np.searchsorted(starts, r, side="right") - 1)               # the file this row range belongs to
            while frames[f] == 0 or r >= starts[f + 1]:
                f += 1
            assert int(offs[f]) <= o and o + l <= int(offs[f]) + int(lens[f])   # (a)
            assert r == row                                                      # (b)
            assert (o - int(offs[f])) % hop == 0 and (o - int(offs[f])) // hop == r - starts[f]   # a frame boundary
            row += orc.num_frames(int(l), N, hop)                                # (c)
        assert row == starts[-1]

    run()


def test_shard_plan_and_frame_count_for_a_non_power_of_two_size(aus, orc):
    """fft sizes outside the native range are legal (chirp-z path): the pure-host helpers treat them like any other size."""
    N, hop = 1000, 250
    for n in (0, 1, 999, 1000, 1001, 1250, 1251, 123457):
        assert aus.num_frames(n, N, hop) == orc.num_frames(n, N, hop)
    vo, vl, vr, vd = aus.plan_shards([0], [123457], N, hop, 4)
    F = orc.num_frames(123457, N, hop)
    rows = 0
    for o, l, r in zip(vo, vl, vr):
        assert r == rows and o == r * hop                        # contiguous frame ranges, each starting on its first frame
        rows += orc.num_frames(int(l), N, hop)
    assert rows == F and sorted(set(vd)) == [0, 1, 2, 3]
    with pytest.raises(aus.AusError):
        aus.plan_shards([0], [50000], 10000, 2500, 2)            # not a power of two and above 8192: unsupported
