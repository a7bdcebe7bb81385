"""Generate golden fixtures from the reference's OWN Python tests.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It executes /root/reference/tests/FFTDirectSolver/testFftSolver_{1D,2D,3D}.py
unchanged (runpy), then calls *their* functions (build_diag_mat_vec_*,
solve_circulant_system_*, build_C_*) on seeded inputs and stores inputs and
outputs as .npz files next to this script.  The GPU box never sees
/root/reference; tests read only the committed .npz files.
"""
import contextlib
import io
import os
import runpy

import numpy as np

REF = "/root/reference/tests/FFTDirectSolver"
HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    with contextlib.redirect_stdout(io.StringIO()):
        return runpy.run_path(os.path.join(REF, name))


def main():
    m1, m2, m3 = load("testFftSolver_1D.py"), load("testFftSolver_2D.py"), load("testFftSolver_3D.py")

    # --- 1-D, testFftSolver_1D.py:38-42 (size=8, lambda=1, rng(123)) ----------------------
    size, lam = 8, 1.0
    rng = np.random.default_rng(123)
    col = m1["build_circulant_col"](size, lam)
    x_ref = rng.random(size)
    import scipy.linalg as spl
    b = spl.circulant(col) @ x_ref
    x = m1["solve_circulant_system"](col, b)
    np.savez(os.path.join(HERE, "ref_py_1d_n8.npz"), col=col, b=b, x=x, x_ref=x_ref, lam=lam)

    # --- 1-D integer KAT, testFftSolver_1D.c:144-177 (N=4, col [1.5,-0.5], b=i^3) ----------
    col = np.array([1.5, -0.5, 0.0, 0.0])
    b = np.arange(4.0) ** 3
    x = m1["solve_circulant_system"](col, b)
    np.savez(os.path.join(HERE, "ref_c_kat1_n4.npz"), col=col, b=b, x=x)

    # --- 2-D, testFftSolver_2D.py:81-89 (50 x 200, lambda=(3, 0.3)) -----------------------
    n_x, n_y = 50, 200
    lx, ly = 30 * 0.01 / 0.1, 3 * 0.01 / 0.1
    rng = np.random.default_rng(123)
    X_ref = rng.random((n_x, n_y)).flatten()
    C = m2["build_C_2D"](n_x, n_y, lx, ly)
    b = C @ X_ref
    Diag = m2["build_diag_mat_vec_2D"](n_x, n_y, lx, ly)
    X = m2["solve_circulant_system_2D"](Diag, b, n_x, n_y)
    np.savez(os.path.join(HERE, "ref_py_2d_50x200.npz"), n=(n_x, n_y, 1), lam=(lx, ly, 0.0),
             b=b, Diag=Diag, X=X, X_ref=X_ref)

    # --- 2-D integer KAT, testFftSolver_2D.c:275-319 (3 x 2, lambda=(1,1), X_ref=m^3) ------
    n_x, n_y = 3, 2
    X_ref = np.arange(n_x * n_y, dtype=float) ** 3
    C = m2["build_C_2D"](n_x, n_y, 1.0, 1.0)
    b = C @ X_ref
    Diag = m2["build_diag_mat_vec_2D"](n_x, n_y, 1.0, 1.0)
    X = m2["solve_circulant_system_2D"](Diag, b, n_x, n_y)
    np.savez(os.path.join(HERE, "ref_c_kat2_3x2.npz"), n=(n_x, n_y, 1), lam=(1.0, 1.0, 0.0),
             b=b, Diag=Diag, X=X, X_ref=X_ref)

    # --- 3-D, testFftSolver_3D.py:82-93 (10 x 25 x 40, lambda=(0.6, 0.15, 0.02)) ----------
    n_x, n_y, n_z = 10, 25, 40
    lx, ly, lz = 6 * 0.01 / 0.1, 3 * 0.01 / 0.2, 1 * 0.01 / 0.5
    rng = np.random.default_rng(123)
    X_ref = rng.random((n_x, n_y, n_z)).flatten()
    C = m3["build_C_3D"](n_x, n_y, n_z, lx, ly, lz)
    b = C @ X_ref
    Diag = m3["build_diag_mat_vec_3D"](n_x, n_y, n_z, lx, ly, lz)
    X = m3["solve_circulant_system_3D"](Diag, b, n_x, n_y, n_z)
    np.savez(os.path.join(HERE, "ref_py_3d_10x25x40.npz"), n=(n_x, n_y, n_z), lam=(lx, ly, lz),
             b=b, Diag=Diag, X=X, X_ref=X_ref)

    # --- 3-D integer KAT, testFftSolver_3D.c:95-141 (4 x 3 x 2, lambda=1, X_ref=m^3) -------
    n_x, n_y, n_z = 4, 3, 2
    X_ref = np.arange(n_x * n_y * n_z, dtype=float) ** 3
    C = m3["build_C_3D"](n_x, n_y, n_z, 1.0, 1.0, 1.0)
    b = C @ X_ref
    Diag = m3["build_diag_mat_vec_3D"](n_x, n_y, n_z, 1.0, 1.0, 1.0)
    X = m3["solve_circulant_system_3D"](Diag, b, n_x, n_y, n_z)
    np.savez(os.path.join(HERE, "ref_c_kat3_4x3x2.npz"), n=(n_x, n_y, n_z), lam=(1.0, 1.0, 1.0),
             b=b, Diag=Diag, X=X, X_ref=X_ref)

    # --- BASELINE config 0: 32^3 direct solve through the reference's Python functions ------
    # (b is built matrix-free: dense C would be 32768^2.)  Only X.real is stored (|X.imag| max is kept).
    n = 32
    for tag, lam in (("phys", (0.6, 0.15, 0.02)), ("unit", (1.0, 1.0, 1.0))):
        rng = np.random.default_rng(123)
        X_ref = rng.random(n ** 3)
        u = X_ref.reshape(n, n, n)
        b = u.copy()
        for ax, l in zip((2, 1, 0), lam):
            b = b + l * (u - np.roll(u, 1, axis=ax))
        b = b.reshape(-1)
        Diag = m3["build_diag_mat_vec_3D"](n, n, n, *lam)
        X = m3["solve_circulant_system_3D"](Diag, b, n, n, n)
        np.savez_compressed(os.path.join(HERE, f"ref_py_3d_32cube_{tag}.npz"), n=(n, n, n), lam=lam,
                            b=b, X_real=X.real, X_imag_max=np.abs(X.imag).max(),
                            Diag_head=Diag[:64], X_ref=X_ref)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
