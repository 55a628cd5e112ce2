"""CPU restatement of the reference's 3-D bilateral solver (stage 3).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Two restatements of /root/reference/bilateral_solver3d.py:

  ``solve_sparse``  keeps the reference's data structures: hashed 6-D lattice
      coordinates + np.unique (:39-61), 0/1 CSR splat matrix (:66-67), per-dimension
      +-1 neighbour adjacency among EXISTING vertices (:71-81), bistochastisation
      (:107-118), Jacobi-PCG through scipy.sparse.linalg.cg (:128-154).  It is the
      honest CPU baseline (``cpu_baseline.kind == "port"``).
  ``solve_dense``   the same arithmetic on a dense (Gx,Gy,Gz,L) grid with an
      occupancy mask and a matrix-free PCG -- the formulation the
      CUDA kernels implement.

The PCG loop of ``solve_dense`` restates SciPy's ``_isolve/iterative.py::cg`` (scipy
is a pip dependency of the reference via scikit-learn, version unpinned; 1.18.1 in
this image): stop when ||r||_2 < rtol*||b||_2 tested at the top of an iteration, at
most ``cg_maxiter`` updates, atol = 0.

Pinned by tests/golden/bls_*.npz (outputs of the reference's own
``apply_bilateral_solver3d`` run with the SciPy>=1.14 ``tol``->``rtol`` shim).
"""
import numpy as np
import torch
import torch.nn.functional as F
from scipy.sparse import csr_matrix, diags
from scipy.sparse.linalg import cg as scipy_cg

GRID_DEFAULTS = {"sigma_luma": 4, "sigma_chroma": 4, "sigma_spatial": 24}      # :156-160
SOLVER_DEFAULTS = {"lam": 256, "A_diag_min": 1e-5, "cg_tol": 1e-5, "cg_maxiter": 25}  # :162-167

_YUV = np.array([[0.299, 0.587, 0.114], [-0.168736, -0.331264, 0.5], [0.5, -0.418688, -0.081312]])
_YUV_OFF = np.array([0, 128.0, 128.0])
N_LATTICE_DIMS = 6  # x, y, z, luma, u, v -- `blur` uses 2*dim = 12 even for grey input (:96)


def luma_lut(sigma_luma):
    """Bin of every grey level, with the reference's exact float64 expression
    (:20,46): tensordot of (v,v,v) with the YUV matrix, / sigma, astype(int).
    NOT v // sigma (differs at 9-22 of 256 levels)."""
    v = np.arange(256, dtype=np.uint8)
    rgb = np.stack([v, v, v], -1)[None, None]                       # (1,1,256,3)
    y = (np.tensordot(rgb, _YUV, ([3], [1])) + _YUV_OFF.reshape(1, 1, -1))[0, 0, :, 0]
    return (y / sigma_luma).astype(int)


def sobel_confidence(r0_u8):
    """:176-181,233-237 -- fp32 central differences with zero padding on r/255,
    sqrt of the sum of squares, then ``max - c`` in float64."""
    v = r0_u8[None, None].float() / 255.0
    k = torch.tensor([-0.5, 0.0, 0.5])
    gz = F.conv3d(v, k.view(1, 1, 1, 1, 3), padding=(0, 0, 1)) ** 2
    gy = F.conv3d(v, k.view(1, 1, 1, 3, 1), padding=(0, 1, 0)) ** 2
    gx = F.conv3d(v, k.view(1, 1, 3, 1, 1), padding=(1, 0, 0)) ** 2
    c = (gz + gy + gx).sqrt()[0, 0]
    return (c.max() - c).numpy().astype(np.float64)


def _inputs(t, r, c, grid_params, bs_params):
    gp = {**GRID_DEFAULTS, **grid_params}
    bs = {**SOLVER_DEFAULTS, **bs_params}
    shape = tuple(t.shape[-3:])
    tt = t.reshape(shape).double().numpy() if isinstance(t, torch.Tensor) else np.asarray(t, np.float64).reshape(shape)
    r = torch.as_tensor(r)
    conf = sobel_confidence(r[0]) if c is None else torch.as_tensor(c).reshape(shape).double().numpy()
    return gp, bs, shape, tt, r, conf


# ----------------------------------------------------------------------------- sparse port
def _lattice_coords(r_whd3, gp):
    yuv = np.tensordot(r_whd3, _YUV, ([3], [1])) + _YUV_OFF.reshape(1, 1, 1, -1)
    i0, i1, i2 = np.mgrid[:r_whd3.shape[0], :r_whd3.shape[1], :r_whd3.shape[2]]
    ss = gp["sigma_spatial"]
    cols = [(i2 / ss).astype(int), (i1 / ss).astype(int), (i0 / ss).astype(int),
            (yuv[..., 0] / gp["sigma_luma"]).astype(int),
            (yuv[..., 1] / gp["sigma_chroma"]).astype(int),
            (yuv[..., 2] / gp["sigma_chroma"]).astype(int)]
    return np.stack([c.reshape(-1) for c in cols], axis=1)


class SparseLattice:
    def __init__(self, r_whd3, gp):
        coords = _lattice_coords(r_whd3, gp)
        self.npix, self.dim = coords.shape
        weights = 255.0 ** np.arange(self.dim)
        keys = coords @ weights
        uniq, first, inv = np.unique(keys, return_index=True, return_inverse=True)
        verts = coords[first]
        self.nvert = len(uniq)
        self.S = csr_matrix((np.ones(self.npix), (inv, np.arange(self.npix))), shape=(self.nvert, self.npix))
        self.adj = []
        for d in range(self.dim):
            a = csr_matrix((self.nvert, self.nvert))
            for step in (-1.0, 1.0):
                shifted = verts.astype(np.float64).copy()
                shifted[:, d] += step
                nk = shifted @ weights
                pos = np.clip(np.searchsorted(uniq, nk), 0, self.nvert - 1)
                hit = np.flatnonzero(uniq[pos] == nk)
                a = a + csr_matrix((np.ones(len(hit)), (hit, pos[hit])), shape=(self.nvert, self.nvert))
            self.adj.append(a)

    def splat(self, x):
        return self.S @ x

    def slice(self, y):
        return self.S.T @ y

    def blur(self, x):
        out = 2 * self.dim * x
        for a in self.adj:
            out = out + a @ x
        return out


def solve_sparse(t, r, c=None, grid_params={}, bs_params={}):
    """apply_bilateral_solver3d (:211-245), CSR formulation.  t (1,W,H,D) float,
    r (3,W,H,D) uint8 -> float32 (W,H,D)."""
    gp, bs, shape, tt, r, conf = _inputs(t, r, c, grid_params, bs_params)
    lat = SparseLattice(r.permute(1, 2, 3, 0).numpy(), gp)
    x = tt.reshape(-1)
    w = conf.reshape(-1)
    m = lat.splat(np.ones(lat.npix))
    n = np.ones(lat.nvert)
    for _ in range(10):
        n = np.sqrt(n * m / lat.blur(n))
    m = n * lat.blur(n)
    Dn, Dm = diags(n, 0), diags(m, 0)
    w_splat = lat.splat(w)
    A = bs["lam"] * (Dm - Dn @ lat.blur(Dn)) + diags(w_splat, 0)
    b = lat.splat(x * w)
    Minv = diags(1.0 / np.maximum(A.diagonal(), bs["A_diag_min"]), 0)
    with np.errstate(invalid="ignore", divide="ignore"):
        y0 = b / w_splat
    y, _ = scipy_cg(A, b, x0=y0, M=Minv, maxiter=bs["cg_maxiter"], rtol=bs["cg_tol"], atol=0.0)
    out = lat.slice(y).reshape(shape)
    return torch.nan_to_num(torch.from_numpy(out).to(torch.float32)), lat.nvert


# ----------------------------------------------------------------------------- dense grid
class DenseGrid:
    """Grey reference => chroma bins are constants, the lattice is a dense 4-D box
    (SURVEY.md App. D4).  Cells never hit by a voxel are masked out of every blur."""

    def __init__(self, r0_u8, gp):
        W, H, D = r0_u8.shape
        ss = gp["sigma_spatial"]
        lut = luma_lut(gp["sigma_luma"])
        b0, b1, b2 = [(np.arange(n) / ss).astype(int) for n in (W, H, D)]       # same expression as :43-45
        self.dims = (int(b0[-1]) + 1, int(b1[-1]) + 1, int(b2[-1]) + 1, int(lut.max()) + 1)
        i0, i1, i2 = np.meshgrid(b0, b1, b2, indexing="ij")
        lum = lut[r0_u8.numpy()]
        g = self.dims
        self.idx = (((i0 * g[1] + i1) * g[2] + i2) * g[3] + lum).reshape(-1)
        self.ncell = int(np.prod(g))
        self.count = np.bincount(self.idx, minlength=self.ncell).astype(np.float64)
        self.occ = (self.count > 0).reshape(g)

    def splat(self, v):
        return np.bincount(self.idx, weights=v.reshape(-1), minlength=self.ncell).reshape(self.dims)

    def slice(self, y):
        return y.reshape(-1)[self.idx]

    def blur(self, y):
        out = 2 * N_LATTICE_DIMS * y
        for ax in range(4):
            pad = [(0, 0)] * 4
            pad[ax] = (1, 1)
            yp = np.pad(y, pad)
            sl_lo = [slice(None)] * 4
            sl_hi = [slice(None)] * 4
            sl_lo[ax] = slice(0, -2)
            sl_hi[ax] = slice(2, None)
            out = out + yp[tuple(sl_lo)] + yp[tuple(sl_hi)]
        return np.where(self.occ, out, 0.0)


def solve_dense(t, r, c=None, grid_params={}, bs_params={}, return_info=False):
    gp, bs, shape, tt, r, conf = _inputs(t, r, c, grid_params, bs_params)
    g = DenseGrid(r[0], gp)
    occ = g.occ
    m = g.count.reshape(g.dims)
    n = occ.astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        for _ in range(10):
            n = np.where(occ, np.sqrt(n * m / g.blur(n)), 0.0)
        m = n * g.blur(n)
        wbar = g.splat(conf)
        b = g.splat(tt * conf)
        lam = bs["lam"]

        def A(p):
            return lam * (m * p - n * g.blur(n * p)) + wbar * p

        minv = np.where(occ, 1.0 / np.maximum(lam * (m - 2 * N_LATTICE_DIMS * n * n) + wbar, bs["A_diag_min"]), 0.0)
        y = np.where(occ, b / wbar, 0.0)
        # NaN from 0/0 on occupied cells with zero confidence propagates exactly as in scipy
        res = b - A(y)
        tol = bs["cg_tol"] * np.sqrt(np.sum(b * b))
        rho_prev, p, iters = None, None, 0
        for it in range(bs["cg_maxiter"]):
            if np.sqrt(np.sum(res * res)) < tol:
                break
            z = minv * res
            rho = np.sum(res * z)
            p = z.copy() if it == 0 else z + (rho / rho_prev) * p
            q = A(p)
            alpha = rho / np.sum(p * q)
            y = y + alpha * p
            res = res - alpha * q
            rho_prev = rho
            iters += 1
    out = torch.nan_to_num(torch.from_numpy(g.slice(y).reshape(shape)).to(torch.float32))
    if return_info:
        return out, {"nvert": int(occ.sum()), "dims": g.dims, "iters": iters}
    return out


# ----------------------------------------------------------------------------- crop helpers
def crop_pad(tensors, thresh=0.1, pad=0):
    """:183-204 -- bounding box of ``tensors[0] > thresh`` padded and clamped."""
    first = tensors[0]
    nz = torch.nonzero(first > thresh)
    lo = torch.clamp(nz.min(dim=0).values[-3:] - pad, 0, None)
    hi = torch.minimum(nz.max(dim=0).values[-3:] + pad + 1, torch.tensor(first.shape[-3:]))
    return [s[..., lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] for s in tensors], (lo, hi)


def write_crop_into(full, crop, lohi):
    """:206-209"""
    lo, hi = lohi
    full[..., lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = crop
    return full
