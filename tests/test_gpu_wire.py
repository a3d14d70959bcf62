"""sqg_load_concordant_wire (chunked upload + k_wire_decode, squid_b200/csrc/sq_wire.cuh): the resident batch after a wire upload
is the batch the wire was packed from, array for array, and the path on top of it gives the reference's results."""
import numpy as np
import pytest

from squid_b200 import api, synth
from tests import common
from tests.test_cpu_wire import adversarial_batch

pytestmark = pytest.mark.gpu


def _graph(n_ref=3):
    return api.SegmentGraph(api.Config(ReadLen=100), np.full(n_ref, 1 << 30, np.int32), device=0)


@pytest.mark.parametrize("seed,n", [(1, 3000), (2, 70000), (3, 513), (4, 1)])
def test_wire_upload_reproduces_batch(seed, n):
    b = adversarial_batch(seed, n)
    g = _graph()
    for pinned in (False, True):
        w = api.WireBatch(b, pinned=pinned)
        g.load_concordant_wire(w)
        d = g.download_concordant()
        for k, v in b.a.items():
            assert np.array_equal(d.a[k], v), k
    g.close()


def test_wire_upload_after_plain_upload_and_empty():
    g = _graph()
    b = adversarial_batch(9, 2000)
    g.load_concordant(b)
    e = api.RecordBatch({k: np.zeros(1 if k == "blk_off" else 0, dt) for k, dt in api.BATCH_DTYPES.items()})
    g.load_concordant_wire(api.WireBatch(e))
    assert g.stat("n_rec") == 0
    g.load_concordant_wire(api.WireBatch(b))
    d = g.download_concordant()
    assert all(np.array_equal(d.a[k], v) for k, v in b.a.items())
    g.close()


def test_inconsistent_wire_is_reported():
    b = adversarial_batch(5, 3000)
    w = api.WireBatch(b)
    import ctypes as C
    arr = w.arrays()
    listed = set(arr["rec_exc"]["idx"].tolist())
    j = next(i for i in range(b.n_rec) if i not in listed)
    (C.c_uint16 * 1).from_address(w.struct.dpos + 2 * j)[0] = 0xFFFF  # an escape without an entry in rec_exc
    g = _graph()
    g.load_concordant_wire(w)
    with pytest.raises(api.SquidB200Error):
        g.download_concordant()
    g.close()


def test_path_on_wire_upload_matches_reference(ref_oracle, tmp_path):
    cp, hp, conc, chim, info = common.write_case(str(tmp_path), 150000, seed=11, disc_frac=0.02, ref_len=synth.CHR17_LEN)
    ref = ref_oracle.run(cp, hp, str(tmp_path / "ref"))
    case = api.HostCase(cp, hp)
    g = api.SegmentGraph(case.config, case.ref_len, device=0)
    g.load_concordant_wire(api.WireBatch(case.batch, pinned=True))
    nodes = g.BuildNode_STAR(case.chimeric)
    edges = g.BuildEdges()
    got = {"nodes": np.stack([nodes.Chr, nodes.Position, nodes.Length, nodes.Support], axis=1).astype(np.int32), "avgdepth": nodes.AvgDepth,
           "edges": edges.table(), "chim_after_edges": case.chimeric.block_table()}
    common.assert_same(ref, got)
    g.close()
