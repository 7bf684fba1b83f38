"""Host orchestration of fast3r_b200.Fast3R on the CPU: the kernels are replaced by tests/abi_emulator.py (an
executable spec of the C ABI), everything else - view grouping, batch permutes, RNG replay, hook bookkeeping, DPT
wiring, head chunking, result assembly, inference() collation - is the product code.  Compared against the
reference-generated fixtures (tolerance: the emulator rounds operands to bf16 like the kernels do)."""
import os

import numpy as np
import pytest
import torch

from tests.conftest import rel_l2
from tests.golden.synth import synth_state_dict, synth_images

TAGS = ["tiny_b1_n3", "tiny_b2_n2", "tiny_noattnbias", "tiny_fixedidx", "tiny_nolocal_n2", "tiny_single_view",
        "tiny_trainmode"]


@pytest.fixture()
def emulated(monkeypatch):
    import fast3r_b200.model as M
    from tests import abi_emulator
    monkeypatch.setattr(M, "ops", abi_emulator)
    monkeypatch.setattr(M, "_require_cuda", lambda device: None)
    return M


def _model(M, g):
    from fast3r_b200 import tiny_args
    enc, dec, head = tiny_args()
    dec.update(g.get("dec_over", {}))
    head.update(g.get("head_over", {}))
    model = M.Fast3R(enc, dec, head).eval()
    model.load_state_dict(synth_state_dict(g["shapes"], seed=g["weight_seed"]))
    if g.get("train_mode", False):
        model.train()
    return model


@pytest.mark.parametrize("tag", TAGS)
def test_forward_against_reference_fixture(emulated, golden_dir, tag):
    g = torch.load(os.path.join(golden_dir, f"{tag}.pt"))
    model = _model(emulated, g)
    model.set_max_parallel_views_for_head(2)  # exercise head chunking
    imgs = synth_images(g["N"], g["B"], g["H"], g["W"])
    torch.manual_seed(g["rng_seed"])
    with torch.no_grad():
        preds = model([dict(img=im) for im in imgs])
    assert len(preds) == g["N"]
    for p, q in zip(preds, g["preds"]):
        assert sorted(p) == sorted(q)
        for k in q:
            assert p[k].shape == q[k].shape and p[k].dtype == torch.float32
    for k in g["preds"][0]:
        a = torch.cat([p[k].flatten() for p in preds])
        b = torch.cat([p[k].flatten() for p in g["preds"]])
        assert rel_l2(a, b) < 2e-2, (tag, k, rel_l2(a, b))


def test_mixed_resolution_against_reference_fixture(emulated, golden_dir):
    g = torch.load(os.path.join(golden_dir, "tiny_mixed_res.pt"))
    model = _model(emulated, g)
    imgs = [synth_images(1, g["B"], h, w, seed0=1234 + i)[0] for i, (h, w) in enumerate(g["sizes"])]
    torch.manual_seed(g["rng_seed"])
    preds = model([dict(img=im) for im in imgs])
    for i, q in enumerate(g["preds"]):
        for k in q:
            assert preds[i][k].shape == q[k].shape
            assert rel_l2(preds[i][k], q[k]) < 3e-2, (i, k, rel_l2(preds[i][k], q[k]))


def test_inference_api_structure(emulated, golden_dir):
    from fast3r_b200 import inference
    g = torch.load(os.path.join(golden_dir, "tiny_b1_n3.pt"))
    model = _model(emulated, g)
    imgs = synth_images(g["N"], g["B"], g["H"], g["W"])
    views = [dict(img=im, true_shape=np.int32([[g["H"], g["W"]]]), idx=i, instance=str(i), dataset="synthetic",
                  label=f"v{i}") for i, im in enumerate(imgs)]
    torch.manual_seed(g["rng_seed"])
    res = inference(views, model, torch.device("cpu"), dtype=torch.bfloat16, verbose=False)
    assert sorted(res.keys()) == g["inference_keys"]
    assert sorted(res["views"][0].keys()) == g["inference_view_keys"]
    for p, q in zip(res["preds"], g["preds"]):
        for k in q:
            assert rel_l2(p[k], q[k]) < 2e-2


# ------------------------------------------------------------------ parity path (precision="fp32") host wiring
PARITY_TOL = 1e-3  # north star: pointmaps within 1e-3 rel-L2 of the reference's fp32 path


@pytest.mark.parametrize("tag", ["tiny_b1_n3", "tiny_b2_n2", "tiny_nolocal_n2", "tiny_trainmode"])
def test_parity_path_against_reference_fixture(emulated, golden_dir, tag):
    """precision="fp32": fp32 activations, hi/lo-split bf16 products (emulated here with the same split arithmetic)."""
    g = torch.load(os.path.join(golden_dir, f"{tag}.pt"))
    model = _model(emulated, g).set_precision("fp32")
    model.set_max_parallel_views_for_head(2)
    imgs = synth_images(g["N"], g["B"], g["H"], g["W"])
    torch.manual_seed(g["rng_seed"])
    with torch.no_grad():
        preds = model([dict(img=im) for im in imgs])
    for k in g["preds"][0]:
        a = torch.cat([p[k].flatten() for p in preds])
        b = torch.cat([p[k].flatten() for p in g["preds"]])
        assert preds[0][k].dtype == torch.float32
        assert rel_l2(a, b) < PARITY_TOL, (tag, k, rel_l2(a, b))


def test_parity_path_mixed_resolution(emulated, golden_dir):
    g = torch.load(os.path.join(golden_dir, "tiny_mixed_res.pt"))
    model = _model(emulated, g).set_precision("fp32")
    imgs = [synth_images(1, g["B"], h, w, seed0=1234 + i)[0] for i, (h, w) in enumerate(g["sizes"])]
    torch.manual_seed(g["rng_seed"])
    preds = model([dict(img=im) for im in imgs])
    for i, q in enumerate(g["preds"]):
        for k in q:
            assert rel_l2(preds[i][k], q[k]) < PARITY_TOL, (i, k, rel_l2(preds[i][k], q[k]))


def test_inference_dtype_selects_precision(emulated, golden_dir):
    """inference(dtype="32") and dtype=torch.float32 run the parity path, torch.bfloat16 the fast path
    (reference: fast3r/dust3r/inference_multiview.py:41-49) and the model's own setting is restored afterwards."""
    from fast3r_b200 import inference
    from fast3r_b200.inference import precision_of
    assert precision_of("32") == "fp32" and precision_of(torch.float32) == "fp32"
    assert precision_of(torch.bfloat16) == "bf16" and precision_of("bf16") == "bf16" and precision_of(None) == "bf16"
    g = torch.load(os.path.join(golden_dir, "tiny_nolocal_n2.pt"))
    model = _model(emulated, g)
    imgs = synth_images(g["N"], g["B"], g["H"], g["W"])
    mk = lambda: [dict(img=im, true_shape=np.int32([[g["H"], g["W"]]]), idx=i, instance=str(i))  # noqa: E731
                  for i, im in enumerate(imgs)]
    err = {}
    for dt in ("32", torch.bfloat16):
        torch.manual_seed(g["rng_seed"])
        res = inference(mk(), model, torch.device("cpu"), dtype=dt, verbose=False)
        err[dt] = max(rel_l2(p[k], q[k]) for p, q in zip(res["preds"], g["preds"]) for k in q)
        assert model.precision == "bf16"
    assert err["32"] < PARITY_TOL < err[torch.bfloat16], err


def test_emulated_partial_merge_equals_attention():
    """Spec of f3r_attention_partial / f3r_attention_merge: key ranges attended separately and merged by log-sum-exp
    equal the attention over all keys (what parallel.KVExchange relies on)."""
    from tests import abi_emulator as E
    heads, sq, chunk, world, rank = 2, 70, 90, 3, 1
    D, skv = heads * 64, chunk * world
    g = torch.Generator().manual_seed(3)
    q, kv = torch.randn(sq, D, generator=g).bfloat16(), torch.randn(skv, 2 * D, generator=g).bfloat16()
    lo, hi = rank * chunk, (rank + 1) * chunk
    ranges, splits = [(lo, chunk), (0, lo), (hi, skv - hi)], [1, 1, 1]
    part_o, part_lse = torch.zeros(3, sq, D), torch.zeros(3, 1, heads, sq)
    for i, ((row0, n), ns) in enumerate(zip(ranges, splits)):
        E.attention_partial(q, kv, part_o, part_lse, part_base=i, n_split=ns, batch=1, heads=heads, sq=sq,
                            kv_rows_total=skv, kv_row0=row0, skv=n, scale=0.2)
    out, ref = torch.zeros(sq, D).bfloat16(), torch.zeros(sq, D).bfloat16()
    E.attention_merge(part_o, part_lse, 3, out, batch=1, heads=heads, sq=sq)
    E.attention(q, kv, ref, batch=1, heads=heads, sq=sq, skv=skv, scale=0.2)
    assert rel_l2(out.float(), ref.float()) < 6e-3


# ------------------------------------------------------------------ sequence parallel over gloo (2 ranks, CPU)
def _sp_worker(rank, world, port, tag, golden_dir, ret, seed_skew=0):
    import torch.distributed as dist
    import fast3r_b200.model as M
    from tests import abi_emulator
    from fast3r_b200.parallel import enable_sequence_parallel
    M.ops = abi_emulator
    M._require_cuda = lambda device: None
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.load(os.path.join(golden_dir, f"{tag}.pt"))
    model = _model(M, g)
    model.image_id_rank_offset = 0
    imgs = synth_images(g["N"], g["B"], g["H"], g["W"])
    views = [dict(img=im) for im in imgs]
    torch.manual_seed(g["rng_seed"])
    ref = model(views)                                   # un-sharded forward in this process
    enable_sequence_parallel(model, gather_preds=True)
    # seed_skew != 0: ranks > 0 hold a DIFFERENT CPU RNG state; the ids of rank 0 must still be used everywhere
    torch.manual_seed(g["rng_seed"] + seed_skew * rank)
    out = model(views)                                   # sharded over the 2 gloo ranks
    worst = max(rel_l2(torch.cat([p[k].flatten() for p in out]), torch.cat([p[k].flatten() for p in ref]))
                for k in ref[0])
    fix = max(rel_l2(torch.cat([p[k].flatten() for p in out]), torch.cat([p[k].flatten() for p in g["preds"]]))
              for k in g["preds"][0])
    ret[rank] = (worst, fix)
    dist.destroy_process_group()


@pytest.mark.parametrize("tag,seed_skew", [("tiny_b1_n3", 0), ("tiny_b2_n2", 0), ("tiny_b1_n3", 1000)])
def test_sequence_parallel_forward_over_gloo(golden_dir, tag, seed_skew):
    """2-rank sharded forward vs the un-sharded forward of the same process and vs the reference fixture.  (On the
    GPU the two are bit-identical, tools/sp_check.py; the torch-CPU emulator's matmuls round differently for different
    shapes, so a small tolerance is used here.)"""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_sp_worker, args=(r, 2, port, tag, golden_dir, ret, seed_skew)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    assert set(ret.keys()) == {0, 1}, dict(ret)
    for r in (0, 1):
        worst, fix = ret[r]
        assert worst < 1e-2, (r, worst)
        assert fix < 2e-2, (r, fix)
