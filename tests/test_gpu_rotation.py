"""SURVEY.md 8f-4 slice: dFSRotation (slip boundaries, src/fs/interface/fsrot.c) through the C ABI against a numpy statement of the
semantics the reference documents (fsrot.c:4-37): forward v <- R v with the first ns strong components enforced, reverse their
transpose; rotating forward then back restores the free components; upstream's argument checks."""
import numpy as np
import pytest

from dohp_b200 import lib

pytestmark = pytest.mark.gpu


def _numpy_rotation(v, bs, idx, rmat, ns, strong, reverse, inhomog):
    v = v.copy().reshape(-1, bs)
    s = 0
    for i, node in enumerate(idx):
        R = rmat[i].reshape(bs, bs)
        sv = strong[s:s + ns[i]] if inhomog else np.zeros(ns[i])
        s += ns[i]
        if not reverse:
            out = R @ v[node]
            out[:ns[i]] = sv
        else:
            t = v[node].copy()
            t[:ns[i]] = sv
            out = R.T @ t
        v[node] = out
    return v.reshape(-1)


@pytest.mark.parametrize("bs", [2, 3])
def test_rotation_forward_reverse(bs):
    import torch
    rng = np.random.default_rng(bs)
    ctx = lib.Context(0, stream=torch.cuda.current_stream().cuda_stream)
    mesh = lib.BoxMesh((3, 2, 2), 2)
    space = lib.Space.from_mesh(ctx, mesh, bs=bs)
    nloc = mesh.nc + mesh.ngh
    idx = rng.choice(nloc, size=17, replace=False).astype(np.int32)
    rmat = np.stack([np.linalg.qr(rng.standard_normal((bs, bs)))[0] for _ in idx])
    ns = rng.integers(0, bs + 1, size=len(idx)).astype(np.int32)
    strong = rng.standard_normal(int(ns.sum()))
    v0 = rng.standard_normal(nloc * bs)
    for reverse in (False, True):
        for hmode, inhomog in ((lib.HOMOGENEOUS, False), (lib.INHOMOGENEOUS, True)):
            got = space.rotation_apply(idx, rmat, ns, v0.copy(), strong=strong if inhomog else None, reverse=reverse, hmode=hmode)
            ref = _numpy_rotation(v0, bs, idx, rmat, ns, strong, reverse, inhomog)
            assert np.abs(got - ref).max() <= 1e-14 * np.abs(ref).max()
    # forward (inhomogeneous) then reverse (inhomogeneous) is the identity on vectors whose strong components already hold the values
    w = space.rotation_apply(idx, rmat, ns, v0.copy(), strong=strong, reverse=False, hmode=lib.INHOMOGENEOUS)
    back = space.rotation_apply(idx, rmat, ns, w.copy(), strong=strong, reverse=True, hmode=lib.INHOMOGENEOUS)
    w2 = space.rotation_apply(idx, rmat, ns, back.copy(), strong=strong, reverse=False, hmode=lib.INHOMOGENEOUS)
    assert np.abs(w2 - w).max() <= 1e-13
    # a torch device vector is rotated in place
    tv = torch.from_numpy(v0.copy()).cuda()
    space.rotation_apply(idx, rmat, ns, tv, reverse=False, hmode=lib.HOMOGENEOUS)
    ctx.synchronize()
    assert np.abs(tv.cpu().numpy() - _numpy_rotation(v0, bs, idx, rmat, ns, strong, False, False)).max() <= 1e-14 * np.abs(v0).max()
    with pytest.raises(lib.DohpError):                      # fsrot.c:62
        space.rotation_apply(idx, rmat, np.full(len(idx), bs + 1, dtype=np.int32), v0.copy())
    ctx.close()
