"""vertical_profiles_batch / compute_wind_fields_batch: row by row bitwise equal to the scalar
restatements, which tests/test_oracle.py pins bitwise against the reference's pbl_model.vertical_profiles
(src/bldfm/pbl_model.py:58-204, src/bldfm/utils.py:7-27)."""
import numpy as np
import pytest

from bldfm_b200.pbl_model import compute_wind_fields_batch, vertical_profiles, vertical_profiles_batch
from bldfm_b200.utils import compute_wind_fields


def _met(B, seed):
    rng = np.random.default_rng(seed)
    ws, wd = rng.uniform(1, 8, B), rng.uniform(0, 360, B)
    ust = rng.uniform(0.1, 0.8, B)
    mol = np.where(rng.random(B) < 0.5, -1.0, 1.0) * rng.uniform(20, 2000, B)
    return ws, wd, ust, mol, rng


def test_wind_fields_batch_bitwise():
    ws, wd, _, _, _ = _met(500, 0)
    um, vm = compute_wind_fields_batch(ws, wd)
    for i in range(len(ws)):
        a, b = compute_wind_fields(ws[i], wd[i])
        assert a == um[i] and b == vm[i]


@pytest.mark.parametrize("closure", ["MOST", "MOSTM", "CONSTANT", "OAAHOC"])
@pytest.mark.parametrize("given", ["ustar", "z0"])
def test_profiles_batch_bitwise_rows(closure, given):
    if closure == "OAAHOC" and given == "z0":
        pytest.skip("OAAHOC derives z0 from ustar and tke")
    B = 400
    ws, wd, ust, mol, rng = _met(B, 3)
    um, vm = compute_wind_fields_batch(ws, wd)
    kw = dict(ustar=ust) if given == "ustar" else dict(z0=rng.uniform(0.01, 0.5, B))
    if closure == "OAAHOC":
        kw["tke"] = rng.uniform(0.5, 2.0, B)
    pb = vertical_profiles_batch(32, 10.0, (um, vm), mol=mol, closure=closure, **kw)
    assert len(pb) == B
    lens = set()
    with np.errstate(all="ignore"):
        for i in range(B):
            z, prof = vertical_profiles(32, 10.0, (um[i], vm[i]), mol=mol[i], closure=closure,
                                        **{k: v[i] for k, v in kw.items()})
            zb, profb = pb.row(i)
            lens.add(len(z))
            assert len(z) == len(zb) and np.array_equal(z, zb), i
            for a, b in zip(prof, profb):
                assert np.array_equal(np.ravel(a), b, equal_nan=True), i
            # padding beyond the row's levels is zero
            assert not pb.buf[i, :, len(z):].any()
    assert len(lens) > 1          # ragged batch: the number of levels depends on z0 (pbl_model.py:127)


def test_profiles_batch_scalars_and_heights():
    um, vm = np.array([-3.0, 2.0, 0.5]), np.array([-4.0, 1.0, -6.0])
    zm = np.array([10.0, 5.0, 20.0])
    pb = vertical_profiles_batch(16, zm, (um, vm), ustar=0.4, mol=-50.0)
    for i in range(3):
        z, prof = vertical_profiles(16, zm[i], (um[i], vm[i]), ustar=0.4, mol=-50.0)
        zb, profb = pb.row(i)
        assert np.array_equal(z, zb) and all(np.array_equal(a, b) for a, b in zip(prof, profb))
    with pytest.raises(ValueError, match="Either z0 or ustar"):
        vertical_profiles_batch(16, 10.0, (um, vm), ustar=0.4, z0=0.1)
    with pytest.raises(ValueError, match="Invalid closure type"):
        vertical_profiles_batch(16, 10.0, (um, vm), ustar=0.4, closure="nope")
