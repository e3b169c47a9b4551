"""Generates tests/golden/*.npz by running the reference's own matrix_fill_full.

Run in the build container (needs /root/reference):
    oracle/ref_build/build.sh && python tests/golden/make_golden.py
Each fixture stores the state vectors handed to the reference and what it returned: the MSR
graph ``ija``, ``First_Unknown``, the Dirichlet table, ``a`` (ams->val) and ``resid_vector``.
The problem definition itself is rebuilt from ``tests/cases.py`` by name.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_driver  # noqa: E402
from tests.cases import GOLDEN_CASES, case_state  # noqa: E402


def main(names):
    assert ref_driver.ref_available(), "build oracle/_ref first (oracle/ref_build/build.sh)"
    for name in names:
        p, kw, st = case_state(name)
        mp = ref_driver.run_map(p)
        res = ref_driver.run_fill(p, [st], delta_t=kw.get("delta_t", 0.0), theta=kw.get("theta", 0.0),
                                  time=kw.get("time", 0.0))[0]
        assert res["err"] == 0
        out = {"ija": mp["ija"][:-1], "first_unknown": mp["first_unknown"], "dbc": mp["dbc"],
               "x_dirichlet": mp["x_dirichlet"], "inter_mask": mp["inter_mask"], "a": res["a"],
               "resid": res["resid"], "h_elem_avg": res["h_elem_avg"], "U_norm": res["U_norm"]}
        for k, v in st.items():
            out["state_" + k] = v
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: unknowns={len(st['x'])} nnz_plus={mp['nnz_plus']} |a|max={np.abs(res['a']).max():.4g} "
              f"|r|max={np.abs(res['resid']).max():.4g} -> {os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main(sys.argv[1:] or GOLDEN_CASES)
