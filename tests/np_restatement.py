"""Second, independent restatement of the reference LBM step in vectorised numpy f32.

Purpose: the reference has no golden vectors, so the C oracle is pinned by
agreement with this separately written restatement of the same WGSL sources
(collide_stream.wgsl:25-88, boundary.wgsl:3-35, init.wgsl:19-63, layout_and_fn.wgsl:38-51)
plus the derived known answers of SURVEY.md §8c.  numpy evaluates each f32 ufunc with one
rounding and never contracts to FMA, like the oracle.  Test infrastructure only.
"""
import numpy as np

F = np.float32


def uniform_arrays(u):
    e = np.array([[u.e_w_max[i][0], u.e_w_max[i][1]] for i in range(9)], np.float32)
    w = np.array([u.e_w_max[i][2] for i in range(9)], np.float32)
    mx = np.array([u.e_w_max[i][3] for i in range(9)], np.float32)
    inv = np.array([u.inversed_direction[i][0] for i in range(9)], np.int64)
    return e, w, mx, inv


def init(u, material, block_iter):
    """init.wgsl: returns buf0, buf1 (9,ny,nx) and the reset mask of accelerate cells."""
    e, w, mx, inv = uniform_arrays(u)
    ny, nx = material.shape
    solid = (material == 2) | (material == 4)
    b0 = np.zeros((9, ny, nx), np.float32)
    b1 = np.zeros((9, ny, nx), np.float32)
    for i in range(9):
        b0[i] = np.where(solid, F(0), w[i])
    if u.fluid_ty == 0:
        temp = F(w[3] * F(0.5))
        b0[1] = np.where(solid, F(0), F(w[1] + temp))
        b0[3] = np.where(solid, F(0), temp)
        b1[1] = b0[1]
        b1[3] = b0[3]
    acc_reset = ((material == 3) | (material == 6)) & (block_iter > 0)
    return b0, b1, acc_reset


def collide_stream(u, rd, wr, material, block_iter, vx_info, vy_info):
    """One collide_stream pass. Mutates wr, material, block_iter in place; returns (ux,uy,rho)."""
    e, w, mx, inv = uniform_arrays(u)
    omega = F(u.omega)
    solid = (material == 2) | (material == 4)
    acc = (material == 3) | (material == 6)
    f = []
    for i in range(9):
        # f_i(x,y) = rd[i][y - e_y, x - e_x] with periodic wrap
        f.append(np.roll(rd[i], shift=(int(e[i, 1]), int(e[i, 0])), axis=(0, 1)))
    rho = np.zeros_like(f[0])
    ux = np.zeros_like(f[0])
    uy = np.zeros_like(f[0])
    for i in range(9):
        rho = rho + f[i]
        ux = ux + e[i, 0] * f[i]
        uy = uy + e[i, 1] * f[i]
    rho = np.minimum(np.maximum(rho, F(0.8)), F(1.2))
    with np.errstate(all="ignore"):
        ux = ux / rho
        uy = uy / rho
        # accelerate cells: countdown first, then override
        dec = acc & (block_iter > 0)
        block_iter[dec] -= 1
        flip = dec & (block_iter == 0)
        material[flip] = 1
        fx = np.where(acc, vx_info, F(0)).astype(np.float32)
        fy = np.where(acc, vy_info, F(0)).astype(np.float32)
        ux = np.where(acc, fx * F(0.5) / rho, ux)
        uy = np.where(acc, fy * F(0.5) / rho, uy)
        usqr = F(1.5) * (ux * ux + uy * uy)
        for i in range(9):
            eu = e[i, 0] * ux + e[i, 1] * uy
            feq = rho * w[i] * (F(1.0) + F(3.0) * eu + F(4.5) * (eu * eu) - usqr)
            Fi = np.where(acc, w[i] * F(3.0) * (e[i, 0] * fx + e[i, 1] * fy), F(0))
            t = f[i] - omega * (f[i] - feq) + Fi
            t = np.where(t > mx[i], mx[i], np.where(t < F(0), F(0), t))
            wr[i] = np.where(solid, wr[i], t)
    ux = np.where(solid, F(0), ux)
    uy = np.where(solid, F(0), uy)
    rho_out = np.where(solid, F(0), rho)
    return ux.astype(np.float32), uy.astype(np.float32), rho_out.astype(np.float32)


def boundary(u, wr, material):
    e, w, mx, inv = uniform_arrays(u)
    ny, nx = material.shape
    solid = (material == 2) | (material == 4)
    ys, xs = np.nonzero(solid)
    for i in range(9):
        qx = xs - int(e[i, 0])
        qy = ys - int(e[i, 1])
        ok = (qx > 0) & (qy > 0) & (qx < nx - 1) & (qy < ny - 1)
        sx, sy, tx, ty = qx[ok], qy[ok], xs[ok], ys[ok]
        val = wr[i][sy, sx].copy()
        wr[inv[i]][ty, tx] = val
        wr[i][sy, sx] = F(0)


class NpSim:
    def __init__(self, nx, ny, info, u):
        self.u = u
        self.nx, self.ny = nx, ny
        info = info.reshape(ny, nx)
        self.material = info["material"].copy()
        self.block_iter = info["block_iter"].copy()
        self.vx = info["vx"].copy()
        self.vy = info["vy"].copy()
        b0, b1, acc_reset = init(u, self.material, self.block_iter)
        self.material[acc_reset] = 1
        self.block_iter[acc_reset] = 0
        self.vx[acc_reset] = 0
        self.vy[acc_reset] = 0
        self.buf = [b0, b1]
        self.swap = 0
        self.macro = None

    def step(self, n=1):
        for _ in range(n):
            rd, wr = self.buf[self.swap], self.buf[1 - self.swap]
            self.macro = collide_stream(self.u, rd, wr, self.material, self.block_iter, self.vx, self.vy)
            boundary(self.u, wr, self.material)
            self.swap ^= 1

    @property
    def current(self):
        return self.buf[self.swap]
