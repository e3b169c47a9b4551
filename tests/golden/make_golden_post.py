"""Generates tests/golden/post_c1_quad9_ns.npz: the reference's row_sum_scaling_scale + Loo/L1/L2 norms applied to
the system its matrix_fill_full assembled for case c1_quad9_ns (needs /root/reference and oracle/_ref)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_driver  # noqa: E402
from tests.cases import case_state  # noqa: E402

p, kw, st = case_state("c1_quad9_ns")
res = ref_driver.run_fill(p, [st], post=True)[0]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "post_c1_quad9_ns.npz"), norms=res["post_norms"],
                    scale=res["post_scale"], a=res["post_a"], resid=res["post_resid"])
print("norms", res["post_norms"], "scale range", res["post_scale"].min(), res["post_scale"].max())
