"""Regenerate the golden fixtures from the reference's shipped data (run in the build container,
where /root/reference exists):   python tests/golden/make_golden.py

fixture_mesh.npz      data/mesh.xml + mesh_facet_region.xml + mesh_physical_region.xml as arrays
                      (the reference's only golden inputs)
fixture_expected.npz  what the oracle derives from them and the KATs pin: canonical CSR pattern,
                      Dirichlet dofs of TestHeatTransfer.json, the exact discrete answer 350-2.5z
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import fem_oracle as fo  # noqa: E402

REF = "/root/reference/data/"


def main():
    coords, cells = fo.read_dolfin_xml_mesh(REF + "mesh.xml")
    fdim, ftags = fo.read_mesh_function_xml(REF + "mesh_facet_region.xml")
    cdim, ctags = fo.read_mesh_function_xml(REF + "mesh_physical_region.xml")
    assert fdim == 2 and cdim == 3
    np.savez_compressed(os.path.join(HERE, "fixture_mesh.npz"), coords=coords, cells=cells,
                        facet_tags=ftags.astype(np.int32), cell_tags=ctags.astype(np.int32))
    settings = json.load(open(REF + "TestHeatTransfer.json"))
    with open(os.path.join(HERE, "TestHeatTransfer.json"), "w") as f:
        json.dump(settings, f, indent=1)
    facets, cf, count = fo.facet_table(cells)
    rp, ci = fo.csr_pattern(cells, coords.shape[0])
    d1 = np.unique(facets[ftags == 1])
    d2 = np.unique(facets[ftags == 2])
    A, b = fo.heat_system(coords, cells, 20.0, [(d1, 350.0), (d2, 300.0)])
    x = fo.solve_direct(A, b)
    np.savez_compressed(os.path.join(HERE, "fixture_expected.npz"), row_ptr=rp, col_idx=ci, dofs_tag1=d1, dofs_tag2=d2,
                        solution=x, analytic=350.0 - 2.5 * coords[:, 2], n_exterior_facets=int((count == 1).sum()))
    print("fixture:", coords.shape, cells.shape, "nnz", ci.size, "bc", d1.size + d2.size,
          "err vs analytic", fo.relative_l2(x, 350.0 - 2.5 * coords[:, 2]))


if __name__ == "__main__":
    main()
