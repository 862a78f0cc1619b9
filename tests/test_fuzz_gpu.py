"""Seeded random-shape sweeps against the oracle (the sweeps that used to live under
profiles/ and never ran in the suite): ragged N, K = 1 ... 520 incl. the column-chunk
path, scalar / vector operands, x data / autodiff (full and factored d_x), propto;
and the categorical family incl. d_x, 41-64 and more than 64 classes, the matrix
product and its reverse sweep.  Sized to finish within a minute on a B200."""
import pytest

pytestmark = pytest.mark.gpu


def test_fuzz_memory_bound_families(gpu):
    from tests import fuzz_glm
    bad = fuzz_glm.run(n_cases=150, seed=7)
    assert not bad, "\n".join(f"{t}: {e}" for t, e in bad)


def test_fuzz_categorical_family(gpu):
    from tests import fuzz_categorical
    bad = fuzz_categorical.run(n_cases=70, seed=2024)
    assert not bad, "\n".join(f"{t}: {e}" for t, e in bad)
