"""Generates tests/golden/ref_blocked_layout.npz by calling THE REFERENCE'S OWN host container
Buffer3D<float> (oracle/_ref/libref_buffer3d.so = /root/reference/src/include/fluid_buffer3D.h behind
oracle/ref_buffer3d_wrapper.cpp).  CPU only:  python tests/golden/make_golden_blocked.py
For each shape the file holds the container's raw storage after b(i,j,k) = 1 + i + nx*(j + ny*k)."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SHAPES = [(13, 9, 20), (16, 8, 8), (17, 16, 9), (5, 3, 2)]   # (nx, ny, nz): ragged, exact, face-sized, tiny


def main():
    L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_buffer3d.so"))
    L.ref_b3d_physical_n.restype = C.c_long
    F = C.POINTER(C.c_float)
    store = {"shapes": np.array(SHAPES)}
    for nx, ny, nz in SHAPES:
        n = L.ref_b3d_physical_n(nx, ny, nz)
        lin = (np.arange(nx * ny * nz, dtype=np.float32) + 1).reshape(nz, ny, nx)
        out = np.zeros(n, np.float32)
        L.ref_b3d_from_linear(out.ctypes.data_as(F), lin.ctypes.data_as(F), nx, ny, nz)
        back = np.zeros_like(lin)
        L.ref_b3d_to_linear(out.ctypes.data_as(F), back.ctypes.data_as(F), nx, ny, nz)
        assert np.array_equal(back, lin)
        store[f"blocked_{nx}x{ny}x{nz}"] = out
    np.savez_compressed(os.path.join(HERE, "ref_blocked_layout.npz"), **store)
    print("wrote", len(SHAPES), "shapes")


if __name__ == "__main__":
    main()
