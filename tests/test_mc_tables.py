"""CPU: the marching-cubes tables are typed in (none exist on this box) and must be
self-validating; the product copy must equal the oracle copy."""

import os
import re

import importlib.util
import os

_spec = importlib.util.spec_from_file_location(
    "mc_tables", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "mc_tables.py"))
M = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(M)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tables_pass_structural_validation():
    assert M.validate()
    assert sum(M.num_tris_table()) == 820  # total triangles of the classic 256-case table


def _parse_header(path):
    txt = open(path).read()
    def arr(name):
        body = txt[txt.index(name):]
        body = body[body.index("{") + 1:body.index("};")]
        return [int(t, 0) for t in re.findall(r"-?0x[0-9a-fA-F]+|-?\d+", body)]
    return arr("MC_EDGE_TABLE[256]"), arr("MC_TRI_TABLE[256][16]"), arr("MC_NUM_TRIS[256]")


def test_product_and_oracle_headers_match_validated_table():
    flat_tri = [v for row in M.tri_table_padded() for v in row]
    for rel in ("oracle/mc_tables.h", "pyslam_b200/csrc/mc_tables.h"):
        e, t, n = _parse_header(os.path.join(ROOT, rel))
        assert e == M.EDGE_TABLE, rel
        assert t == flat_tri, rel
        assert n == M.num_tris_table(), rel
