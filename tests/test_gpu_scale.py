"""GPU: the full-size BASELINE.json configs C3, C4 and C5 on ONE B200 against digests of the UNMODIFIED reference's
output files (tests/golden/golden_scale.json, generated in the dev container by tests/golden/make_golden_scale.py:
C5 = 8192^3 / 512 partitions took the reference 310 s). The node file is hashed chunk by chunk (9.5 GB at C5)."""
import hashlib
import json
import os

import numpy as np
import pytest

from ooc_svo_builder_b200 import meshgen as mg

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "golden_scale.json")))


@pytest.mark.parametrize("name", ["c3", "c4", "c5"])
def test_full_size_config_matches_reference_digest(name):
    from ooc_svo_builder_b200 import SvoBuilder
    from filesum import filesum_array, add
    gold = GOLD[name]
    mesh = mg.make(gold["config"])
    h = hashlib.sha256()
    flat = mesh.tris.reshape(-1).view(np.uint8)
    for lo in range(0, flat.size, 256 << 20):
        h.update(flat[lo:lo + (256 << 20)].tobytes())
    assert h.hexdigest() == gold["mesh_sha256"], "the generated mesh differs from the golden input"
    sb = SvoBuilder(0)
    try:
        prm = sb.make_params(mesh.length, gold["gridsize"], mesh.payload)
        sb.set_triangles(mesh.tris)
        sb.partition(prm, want_counts=False)
        sb.voxelize()
        nv, nn, nd = sb.build()
        assert (nv, nn, nd) == (gold["n_voxels"], gold["n_nodes"], gold["n_data"])
        assert sb.stats()["n_partitions"] == gold["n_partitions"]
        from ooc_svo_builder_b200 import header_bytes
        assert header_bytes(gold["gridsize"], nn, nd).decode() == gold["header"]
        hn, fs = hashlib.sha256(), (0, 0, 0)
        step = 1 << 25                                       # 805 MB of node records per chunk
        for lo in range(0, nn, step):
            chunk = sb.fetch_nodes(lo, min(step, nn - lo))
            hn.update(chunk.tobytes())
            fs = add(fs, filesum_array(chunk.view(np.uint64), lo * 3))
        assert hn.hexdigest() == gold["nodes_sha256"]
        assert [int(x) for x in fs] == gold["nodes_filesum"]      # the checksum the multi-GPU runs are held to
        hd = hashlib.sha256()
        step = 1 << 24
        for lo in range(0, nd, step):
            hd.update(sb.fetch_data(lo, min(step, nd - lo)).tobytes())
        assert hd.hexdigest() == gold["data_sha256"]
    finally:
        sb.close()
