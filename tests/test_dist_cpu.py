"""N>1 host logic on CPU: world_size 2, gloo backend at 127.0.0.1.
 (1) the bench's max-over-ranks aggregation; (2) the tensor-parallel sharding rules (heads / intermediate rows,
 all-reduce(sum) after o_proj and down_proj) reproduce the full oracle layer."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from aha_b200 import dist_util, synth
        from oracle import nn
        from oracle.qwen3 import Qwen3DecoderLayer, prepare_causal_attention_mask
        from oracle.rope import RoPE
        # (1) aggregation: rank r measured (10 + r) ms for 100 steps
        val, ms = dist_util.aggregate_throughput(100, 10.0 + rank, world, dist)
        assert ms == 10.0 + world - 1 and abs(val - world * 100 / (ms * 1e-3)) < 1e-6
        # (2) TP layer == full layer
        cfg = synth.get_config("qwen3", "tiny")
        w = synth.make_weights("qwen3", cfg, 0)
        p = "model.layers.0."
        x = np.random.default_rng(7).standard_normal((1, 9, cfg["hidden_size"])).astype(np.float32)
        cos, sin = RoPE(cfg["head_dim"], cfg["rope_theta"]).forward(0, 9)
        mask = prepare_causal_attention_mask(1, 9)
        full = Qwen3DecoderLayer(cfg, w, p).forward(x, cos, sin, mask)
        ws, s = dist_util.tp_shard_layer(w, p, cfg, rank, world)
        lc = dict(cfg, num_attention_heads=s["nh_l"], num_key_value_heads=s["nkv_l"])
        layer = Qwen3DecoderLayer(lc, ws, p)
        a = layer.attn.forward(nn.rms_norm(x, layer.ln1, layer.eps), cos, sin, mask)          # partial o_proj output
        t = torch.from_numpy(np.ascontiguousarray(a)); dist.all_reduce(t); x1 = x + t.numpy()  # all-reduce #1
        m = layer.mlp.forward(nn.rms_norm(x1, layer.ln2, layer.eps))                          # partial down_proj output
        t = torch.from_numpy(np.ascontiguousarray(m)); dist.all_reduce(t); x2 = x1 + t.numpy()  # all-reduce #2
        err = float(np.abs(x2 - full).max())
        q.put((rank, err))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_tp_and_aggregation():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(world))
    assert set(res) == {0, 1}
    assert max(res.values()) < 1e-5


def test_tp_slices_cover_everything():
    from aha_b200 import dist_util, synth
    cfg = synth.get_config("qwen3vl", "vl2")["text_config"]
    for world in (1, 2, 4, 8):
        seen_q, seen_m = [], []
        for r in range(world):
            s = dist_util.tp_slices(cfg, r, world)
            seen_q += list(range(*s["q"])); seen_m += list(range(*s["mlp"]))
            assert s["nh_l"] * world == cfg["num_attention_heads"]
        assert seen_q == list(range(cfg["num_attention_heads"] * cfg["head_dim"]))
        assert seen_m == list(range(cfg["intermediate_size"]))
    with pytest.raises(ValueError):
        dist_util.tp_slices(cfg, 0, 3)
