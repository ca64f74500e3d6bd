"""The reference-side binding (pytorch_geometric_b200.plugin) on CPU, against the unmodified reference package
(fixture `tg`): what install() rebinds, that uninstall() restores it, that CPU tensors fall through to the
untouched reference code bit for bit (functions, the lazy gather, the subclass layers), and that the extension
operators (`torch.ops.torch_sparse.*`, `torch.ops.b200mp.*`) infer shapes on meta tensors with the reference's
signatures (test/test_edge_index.py:880-929 shapes).  No engine compute runs here: there is no CPU fallback."""
import pytest
import torch


@pytest.fixture
def plugin(tg):
    from pytorch_geometric_b200 import plugin as P
    yield P
    P.uninstall()


def _graph():
    ei = torch.tensor([[0, 1, 2, 3, 4, 5, 0, 2, 2], [1, 2, 3, 4, 5, 0, 3, 2, 0]])
    return torch.randn(6, 8), ei, torch.tensor([0, 1, 0, 1, 0, 1, 0, 1, 1])


def test_install_rebinds_every_seam_and_uninstall_restores(tg, plugin):
    import torch_geometric.edge_index as tei
    import torch_geometric.nn.aggr.base as aggr_base
    import torch_geometric.typing as T
    from torch_geometric.nn.conv.message_passing import MessagePassing
    from torch_geometric.utils import _scatter, _segment
    orig = (_scatter.scatter, tei._spmm, aggr_base.Aggregation.reduce, MessagePassing._index_select, tg.nn.GCNConv,
            _scatter.torch_scatter, T.WITH_TORCH_SCATTER)
    c = plugin.install(layers=True, flip_flags=True)
    assert c["scatter"] >= 50 and c["softmax"] >= 10 and c["spmm"] >= 10 and c["segment"] >= 1
    assert c["edge_index._spmm"] == c["Aggregation.reduce"] == c["MessagePassing._index_select"] == 1
    assert c["extension_modules"] >= 4 and c["flags"] == 5 and c["layers"] >= 16
    assert _scatter.scatter is not orig[0] and _scatter.scatter.__wrapped__ is orig[0] and aggr_base.scatter is _scatter.scatter
    assert T.WITH_TORCH_SCATTER is True and hasattr(_scatter.torch_scatter, "scatter_max") and hasattr(_segment.torch_scatter, "segment_csr")
    assert issubclass(tg.nn.GCNConv, orig[4]) and tg.nn.GCNConv is not orig[4]
    assert plugin.install() == {}                                   # idempotent
    plugin.uninstall()
    now = (_scatter.scatter, tei._spmm, aggr_base.Aggregation.reduce, MessagePassing._index_select, tg.nn.GCNConv,
           _scatter.torch_scatter, T.WITH_TORCH_SCATTER)
    assert all(a is b for a, b in zip(orig, now))


@pytest.mark.parametrize("flip", [False, True])
def test_cpu_tensors_fall_through_bit_for_bit(tg, plugin, flip):
    from torch_geometric.utils import scatter, segment, softmax
    x, ei, _ = _graph()
    idx = ei[1]
    msg = torch.randn(9, 8)
    ptr = torch.tensor([0, 2, 5, 6, 9])
    base = {r: scatter(msg, idx, 0, 6, r) for r in ("sum", "mean", "min", "max", "mul")}
    base_seg = segment(msg, ptr, "mean")
    base_sm = softmax(msg, idx, num_nodes=6)
    convs = {n: getattr(tg.nn, n)(8, 4) for n in ("GCNConv", "SAGEConv", "GraphConv")}
    base_conv = {n: c(x, ei) for n, c in convs.items()}
    plugin.install(flip_flags=flip)
    from torch_geometric.utils import scatter as s2, segment as g2, softmax as m2
    for r, want in base.items():
        assert torch.equal(s2(msg, idx, 0, 6, r), want), r
    assert torch.equal(g2(msg, ptr, "mean"), base_seg) and torch.equal(m2(msg, idx, num_nodes=6), base_sm)
    for n, c in convs.items():                                      # unmodified reference layers: lazy gather not engaged on CPU
        assert torch.equal(c(x, ei), base_conv[n]), n
    with pytest.raises(ValueError, match="one-dimensional"):        # the reference's own argument errors still surface
        s2(msg, ei, 0, 6, "sum")


def test_subclass_layers_are_the_reference_layers_on_cpu(tg, plugin):
    """`B200GCNConv(GCNConv)` etc.: same parameters / state_dict, CPU forward == the parent's, hooks fall through."""
    from pytorch_geometric_b200.plugin import conv as PC
    x, ei, et = _graph()
    mlp = torch.nn.Linear(8, 4)
    cases = [("GCNConv", (8, 4), {}, ()), ("SAGEConv", (8, 4), {"project": True}, ()), ("GraphConv", (8, 4), {}, ()),
             ("GINConv", (mlp, ), {}, ()), ("GATConv", (8, 4), {"heads": 2}, ()), ("GATv2Conv", (8, 4), {"heads": 2}, ()),
             ("TransformerConv", (8, 4), {"heads": 2, "beta": True}, ()), ("RGCNConv", (8, 4, 2), {"num_bases": 2}, (et, )),
             ("FastRGCNConv", (8, 4, 2), {}, (et, ))]
    for name, args, kw, extra in cases:
        ref = getattr(tg.nn, name)(*args, **kw)
        ours = getattr(PC, PC.LAYERS[name])(*args, **kw)
        assert isinstance(ours, type(ref))
        assert {k: tuple(v.shape) for k, v in ours.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.state_dict().items()}
        ours.load_state_dict(ref.state_dict())
        assert torch.equal(ours(x, ei, *extra), ref(x, ei, *extra)), name
    seen = []
    conv = PC.B200SAGEConv(8, 4)
    conv.register_aggregate_forward_hook(lambda m, i, o: seen.append(tuple(o.shape)))
    conv(x, ei)
    assert seen == [(6, 8)]


def test_extension_operator_signatures_on_meta_tensors(tg, plugin):
    """torch.ops.torch_sparse.spmm_{sum,mean,min,max} with torch_sparse's schemas (edge_index.py:1798-1810) and the
    engine's own torch.library namespace: Meta kernels give shapes / dtypes without a device."""
    plugin.install()
    rowptr = torch.tensor([0, 2, 3, 3, 4], device="meta")
    col = torch.zeros(4, dtype=torch.long, device="meta")
    row = torch.zeros(4, dtype=torch.long, device="meta")
    val = torch.empty(4, device="meta")
    mat = torch.empty(3, 16, device="meta")
    assert torch.ops.torch_sparse.spmm_sum(row, rowptr, col, val, None, None, mat).shape == (4, 16)
    assert torch.ops.torch_sparse.spmm_mean(None, rowptr, col, None, None, None, None, mat).shape == (4, 16)
    out, arg = torch.ops.torch_sparse.spmm_max(rowptr, col, val, mat)
    assert out.shape == arg.shape == (4, 16) and arg.dtype == torch.long
    assert torch.ops.b200mp.spmm_csr(rowptr, col, val, mat, 4, "mean").shape == (4, 16)
    assert torch.ops.b200mp.segment_csr(torch.empty(4, 2, 5, device="meta"), rowptr, "max").shape == (4, 2, 5)
    assert torch.ops.b200mp.scatter_coo(mat, col[:3], 7, "sum").shape == (7, 16)
    ks, perm = torch.ops.b200mp.index_sort(col, 9)
    assert ks.shape == perm.shape == (4, ) and perm.dtype == torch.long
    # the shim modules expose exactly the names the reference calls
    from pytorch_geometric_b200.plugin import shims
    ts, pl = shims.torch_scatter_module(), shims.pyg_lib_module()
    assert all(hasattr(ts, n) for n in ("scatter", "scatter_max", "scatter_min", "segment_csr"))
    assert all(hasattr(pl.ops, n) for n in ("softmax_csr", "index_sort", "segment_matmul", "grouped_matmul"))


def test_lazy_rows_mechanics_on_cpu():
    """The lazy gather object: metadata without data, scale folding, materialisation == index_select (* scale)."""
    from pytorch_geometric_b200.plugin.lazy import LazyRows
    x = torch.randn(5, 3, requires_grad=True)
    idx = torch.tensor([0, 2, 2, 4])
    w = torch.tensor([1.0, 2.0, 3.0, 4.0])
    lazy = LazyRows(x, idx)
    assert lazy.shape == (4, 3) and lazy.dim() == 2 and lazy.dtype == torch.float32 and isinstance(lazy, torch.Tensor)
    scaled = w.view(-1, 1) * lazy
    assert isinstance(scaled, LazyRows) and torch.equal(scaled._scale, w)
    assert isinstance(scaled * w.view(-1, 1), LazyRows)
    dense_ = torch.cat([lazy, lazy], dim=-1)                         # any other op gathers, as the reference would
    assert type(dense_) is torch.Tensor and torch.equal(dense_, torch.cat([x[idx], x[idx]], -1))
    assert torch.equal(scaled.materialise(), w.view(-1, 1) * x[idx])
    (scaled + 0).sum().backward()
    assert torch.allclose(x.grad[2], torch.full((3, ), 5.0))


def test_install_routes_fused_aggregation_and_cpu_falls_through(tg, plugin):
    from torch_geometric.nn.aggr.fused import FusedAggregation
    orig = FusedAggregation.forward
    aggr = FusedAggregation(["sum", "mean", "max", "std"])
    x = torch.randn(12, 3)
    index = torch.tensor([0, 0, 1, 1, 1, 3, 3, 3, 3, 4, 4, 0])
    before = aggr(x, index, dim_size=6)
    counts = plugin.install()
    assert counts["fused_aggregation"] == 1 and FusedAggregation.forward is not orig
    after = aggr(x, index, dim_size=6)             # CPU tensors: untouched reference
    assert all(torch.equal(a, b) for a, b in zip(before, after))
    # a list with 'mul' is not fusable by the engine and must fall through as well
    assert len(FusedAggregation(["sum", "mul"])(x, index, dim_size=6)) == 2
    plugin.uninstall()
    assert FusedAggregation.forward is orig
