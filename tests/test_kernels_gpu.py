"""-m gpu: every CUDA kernel, called through the C ABI, against the CPU oracle on the same seeded inputs."""
import pytest

from tests import kernel_checks as K

pytestmark = pytest.mark.gpu


def _id(case):
    fn, kw = case
    return fn.__name__.replace("check_", "") + "-" + "-".join(f"{k}{v}" for k, v in kw.items())


@pytest.mark.parametrize("case", K.CONV_CASES, ids=_id)
def test_conv_family_matches_oracle(case):
    fn, kw = case
    m = fn(**kw)
    assert m["err"] <= m["tol"], m
    assert m.get("pad_intact", True), f"kernel wrote outside its channel slice: {m}"


@pytest.mark.parametrize("case", K.EW_CASES, ids=_id)
def test_hbm_bound_kernels_match_oracle(case):
    fn, kw = case
    m = fn(**kw)
    assert m["err"] <= m["tol"], m
    assert m.get("pad_intact", True), m


def _fid(c):
    f = c[2]
    return (c[0].__name__.replace("check_", "") + "-" + "-".join(f"{k}{v}" for k, v in c[1].items() if k != "seed")
            + "-" + "-".join(f"{k}{v}" for k, v in f.items()))


@pytest.mark.parametrize("case", K.FORCED_CASES, ids=_fid)
def test_every_tile_width_and_split_k_path(case):
    fn, kw, force = case
    m = K.forced(fn, **force, **kw)
    assert m["err"] <= m["tol"], m
    assert m.get("pad_intact", True), m


@pytest.mark.parametrize("case", K.PAIR_CASES, ids=_fid)
def test_cta_pair_kernels_match_oracle(case):
    """All six cta_group::2 instantiations against the oracle (these were selected automatically for
    every launch with more tiles than SMs but only ever reached by a property test)."""
    fn, kw, force = case
    m = K.forced(fn, **force, **kw)
    assert m["plan"]["pair"] == 1, m
    assert m["err"] <= m["tol"], m
    assert m.get("pad_intact", True), m
