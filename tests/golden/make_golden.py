"""Generates tests/golden/*.npz -- small known-answer vectors for the hot path, independent of the oracle's code.

The reference ships no data files: its tests build their fields in code.  The vectors here
restate those fields and record answers obtained WITHOUT the oracle or the CUDA library:
  * decomp_*.npz     decomposition tables of decomp_info_init (src/decomp_2d.f90:382-490, 1016-1206) for ragged grids,
                     computed by a direct numpy restatement of distribute/partition/prepare_buffer in this file;
  * transpose_*.npz  the four pencils of the test2d index field u(i,j,k) = i + (j-1) nx + (k-1) nx ny
                     (examples/test2d/test2d.f90:92-199): a transpose must turn the X pencil into exactly the
                     Y / Z pencil cut out of the same global array;
  * ramp_dft_*.npz   the examples' field (i/nx)(j/ny)(k/nz) (examples/fft_physical_x/fft_r2c_x.f90:66-75) and its 3-D
                     DFT from the closed form R_n[0] = (n+1)/2, R_n[k] = 1/(exp(-2 pi i k/n) - 1),
                     cross-checked here against numpy's pocketfft.
Run:  python tests/golden/make_golden.py     (numpy only; writes next to this file)"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def distribute(n, p):
    """src/decomp_2d.f90:1070-1105: n/p each, the LAST n%p ranks get one more."""
    base, nu = n // p, n % p
    sz = [base + (1 if i >= p - nu else 0) for i in range(p)]
    st = [sum(sz[:i]) for i in range(p)]
    return st, sz


def decomp(nx, ny, nz, p_row, p_col, rank):
    c1, c2 = rank // p_col, rank % p_col  # MPI_CART_CREATE without reorder (decomp_2d_init_fin.f90:95-123)
    x1s, x1 = distribute(nx, p_row)
    y1s, y1 = distribute(ny, p_row)
    y2s, y2 = distribute(ny, p_col)
    z2s, z2 = distribute(nz, p_col)
    d = dict(xst=(0, y1s[c1], z2s[c2]), xsz=(nx, y1[c1], z2[c2]), yst=(x1s[c1], 0, z2s[c2]), ysz=(x1[c1], ny, z2[c2]),
             zst=(x1s[c1], y2s[c2], 0), zsz=(x1[c1], y2[c2], nz), x1dist=x1, y1dist=y1, y2dist=y2, z2dist=z2)
    d["x1cnts"] = [w * d["xsz"][1] * d["xsz"][2] for w in x1]  # prepare_buffer, decomp_2d.f90:1161-1183
    d["y1cnts"] = [d["ysz"][0] * h * d["ysz"][2] for h in y1]
    d["y2cnts"] = [d["ysz"][0] * h * d["ysz"][2] for h in y2]
    d["z2cnts"] = [d["zsz"][0] * d["zsz"][1] * k for k in z2]
    for n in ("x1", "y1", "y2", "z2"):
        c = d[n + "cnts"]
        d[n + "disp"] = [sum(c[:i]) for i in range(len(c))]
    return d


def cut(glob, d, pencil):
    st, sz = d[("xst", "yst", "zst")[pencil]], d[("xsz", "ysz", "zsz")[pencil]]
    return np.asfortranarray(glob[st[0]:st[0] + sz[0], st[1]:st[1] + sz[1], st[2]:st[2] + sz[2]])


def main():
    # ---- decomposition tables
    for shape, grid in (((17, 13, 11), (2, 2)), ((1024, 1024, 513), (2, 4)), ((257, 512, 512), (2, 4)), ((64, 64, 64), (1, 2))):
        out = {}
        for r in range(grid[0] * grid[1]):
            d = decomp(*shape, *grid, r)
            for k, v in d.items():
                out[f"r{r}_{k}"] = np.array(v, dtype=np.int64)
        np.savez_compressed(os.path.join(HERE, f"decomp_{shape[0]}x{shape[1]}x{shape[2]}_{grid[0]}x{grid[1]}.npz"), **out)
    # ---- test2d index field, ragged grids
    for shape, grid in (((17, 13, 11), (2, 2)), ((9, 24, 16), (3, 2))):
        nx, ny, nz = shape
        i, j, k = np.meshgrid(np.arange(1, nx + 1), np.arange(1, ny + 1), np.arange(1, nz + 1), indexing="ij")
        g = (i + (j - 1) * nx + (k - 1) * nx * ny).astype(np.float64)
        out = {}
        for r in range(grid[0] * grid[1]):
            d = decomp(*shape, *grid, r)
            for p, name in enumerate("xyz"):
                out[f"r{r}_{name}"] = cut(g, d, p)
        np.savez_compressed(os.path.join(HERE, f"transpose_{nx}x{ny}x{nz}_{grid[0]}x{grid[1]}.npz"), **out)
    # ---- ramp field and its analytic DFT
    for shape in ((16, 8, 32), (32, 16, 64)):
        nx, ny, nz = shape

        fx, fy, fz = np.arange(1, nx + 1) / nx, np.arange(1, ny + 1) / ny, np.arange(1, nz + 1) / nz
        field = np.asfortranarray(fx[:, None, None] * fy[None, :, None] * fz[None, None, :])
        # closed form: DFT_k of (i/n), i = 1..n placed at positions 0..n-1:  sum_{m=0}^{n-1} (m+1)/n w^{mk}
        def dft_line(n):
            k = np.arange(n)
            out = np.empty(n, dtype=np.complex128)
            out[0] = (n + 1) / 2.0
            w = np.exp(-2j * np.pi * k[1:] / n)
            out[1:] = 1.0 / (w - 1.0)  # sum_m (m+1) w^{mk} = n / (w^k - 1) for w^n = 1, k != 0; divided by n
            return out
        sx, sy, sz = dft_line(nx), dft_line(ny), dft_line(nz)
        spec = sx[:, None, None] * sy[None, :, None] * sz[None, None, :]
        chk = np.fft.fftn(field)
        err = np.max(np.abs(spec - chk)) / np.max(np.abs(chk))
        assert err < 1e-12, err
        np.savez_compressed(os.path.join(HERE, f"ramp_dft_{nx}x{ny}x{nz}.npz"), field=field, spectrum=np.asfortranarray(spec))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
