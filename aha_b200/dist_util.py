"""Host-side multi-GPU plumbing shared by bench.py and the CPU (gloo) tests: rank layout, the max-over-ranks
timing reduction, and the tensor-parallel slicing rules the C++ loader applies (text_model.cuh: TextModel::load)."""
import numpy as np


def reduce_max(value, dist=None, device=None):
    """max over ranks of a python float (CUDA-event ms / wall seconds)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_throughput(steps, ms_local, world, dist=None, device=None):
    """Whole-job tokens/s for `world` independent replicas each doing `steps` steps: units of all ranks divided by
    the slowest rank's time."""
    ms = reduce_max(ms_local, dist, device)
    return world * steps / (ms * 1e-3), ms


def tp_slices(cfg, rank, world):
    """Row/column ranges of one tensor-parallel rank (heads for q/k/v/o, intermediate rows for gate/up/down)."""
    nh, nkv, hd, I = cfg["num_attention_heads"], cfg["num_key_value_heads"], cfg["head_dim"], cfg["intermediate_size"]
    if nkv % world or I % world:
        raise ValueError("tensor-parallel world must divide num_key_value_heads and intermediate_size")
    g = nh // nkv
    nkv_l, I_l = nkv // world, I // world
    nh_l = nkv_l * g
    return dict(q=(rank * nh_l * hd, (rank + 1) * nh_l * hd), kv=(rank * nkv_l * hd, (rank + 1) * nkv_l * hd),
                o_cols=(rank * nh_l * hd, (rank + 1) * nh_l * hd), mlp=(rank * I_l, (rank + 1) * I_l), nh_l=nh_l, nkv_l=nkv_l)


def tp_shard_layer(w, prefix, cfg, rank, world):
    """Weights of one decoder layer as seen by `rank` (numpy views)."""
    s = tp_slices(cfg, rank, world)
    p = prefix
    out = {
        p + "self_attn.q_proj.weight": w[p + "self_attn.q_proj.weight"][s["q"][0]:s["q"][1]],
        p + "self_attn.k_proj.weight": w[p + "self_attn.k_proj.weight"][s["kv"][0]:s["kv"][1]],
        p + "self_attn.v_proj.weight": w[p + "self_attn.v_proj.weight"][s["kv"][0]:s["kv"][1]],
        p + "self_attn.o_proj.weight": w[p + "self_attn.o_proj.weight"][:, s["o_cols"][0]:s["o_cols"][1]],
        p + "mlp.gate_proj.weight": w[p + "mlp.gate_proj.weight"][s["mlp"][0]:s["mlp"][1]],
        p + "mlp.up_proj.weight": w[p + "mlp.up_proj.weight"][s["mlp"][0]:s["mlp"][1]],
        p + "mlp.down_proj.weight": w[p + "mlp.down_proj.weight"][:, s["mlp"][0]:s["mlp"][1]],
    }
    for n in ("self_attn.q_norm.weight", "self_attn.k_norm.weight", "input_layernorm.weight", "post_attention_layernorm.weight"):
        out[p + n] = w[p + n]
    return out, s
