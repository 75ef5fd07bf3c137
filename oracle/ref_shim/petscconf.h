/* oracle shim: intentionally empty (real-double, 32-bit-index PETSc configuration). */
