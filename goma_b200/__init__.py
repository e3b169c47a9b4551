"""goma_b200 -- B200 (sm_100a) implementation of Goma's matrix_fill assembly hot path.

Layout: ``csrc/`` holds the CUDA kernels and the C ABI declared in
``include/goma_gpu_fill.h``; ``matrix_fill.py`` mirrors the reference's
``matrix_fill_full`` call on top of it; ``problem.py`` / ``mesh.py`` are the host-side
snapshot of the Goma state the path reads; ``dp_comm.py`` is the ghost exchange.
"""
from .mesh import Mesh, box_mesh  # noqa: F401
from .problem import Dirichlet, Problem  # noqa: F401
