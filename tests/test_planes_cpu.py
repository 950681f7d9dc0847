"""Host side of the device read format: the low-quality bit plane and the index list of non-ACGT
bases must say exactly what the quality bytes say (every base-quality test of the reference is
``q < MIN_BASE_QUAL``, read_collector.py:44-46, :123, :281-284)."""
import numpy as np

from unfazed_b200.schema import QUAL_ESCAPE, min_base_qual
from unfazed_b200.synth import SynthConfig, make_dataset


def test_min_base_qual_is_the_integer_threshold():
    assert [min_base_qual(x) for x in (-3, 0, 0.2, 19.0, 19.5, 20, 127.5, 128, 500)] == [0, 0, 1, 19, 20, 20, 128, 128, 128]
    for t in (0.5, 13, 19.01, 20, 40):
        for q in range(0, 94):
            assert (q < t) == (q < min_base_qual(t))


def test_planes_restate_the_quality_bytes():
    ds = make_dataset(SynthConfig(dnms_per_trio=6, seed=3, coverage=8.0))
    t = ds.reads
    q = t.qual
    for thr in (0, 9, 20, 128):
        plane = t.lowq_plane(thr, chunk=1 << 12)
        bits = np.unpackbits(plane, bitorder="little")[: q.shape[0]].astype(bool)
        assert np.array_equal(bits, (q & 0x7F) < thr)
        assert np.array_equal(plane, t.lowq_plane(thr))
    n = t.n_index(chunk=1 << 12)
    assert np.array_equal(n, np.flatnonzero(q & QUAL_ESCAPE))
    assert n.size > 0 and np.all(np.diff(n) > 0)
