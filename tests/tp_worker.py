"""Tensor-parallel check, one process per GPU (launched by tests/test_tp_gpu.py through torch.distributed.run):
every rank builds its shard of the tiny Qwen3 model, prefill + teacher-forced decode logits must match the
single-GPU oracle within 1e-3 on every rank, and all ranks must agree bit-for-bit with each other."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    from aha_b200 import B200Model, nccl_unique_id, synth
    from oracle.qwen3 import Qwen3Model
    for preset in ("tiny", "mid"):
        run(preset, rank, world, local, dist, torch, B200Model, synth, Qwen3Model, nccl_unique_id)
    run_vl(rank, world, local, dist, torch, B200Model, synth, nccl_unique_id)
    dist.destroy_process_group()


def run_vl(rank, world, local, dist, torch, B200Model, synth, nccl_unique_id):
    """Qwen3-VL with three images under tensor parallelism: the ViT is sharded by image over the ranks (embeddings broadcast by
    their owners), the text stack by heads; prefill + decode logits against the single-GPU oracle on every rank."""
    from oracle.qwen3vl import Qwen3VLModel, process_image
    uid = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    cfg = synth.get_config("qwen3vl", "tiny")
    w = synth.make_weights("qwen3vl", cfg, 0)
    m = B200Model("qwen3vl", cfg, w, device=local, max_ctx=1024, max_patches=2048, tp_rank=rank, tp_world=world, tp_unique_id=uid[0])
    o = Qwen3VLModel(cfg, w)
    pvs, grids = zip(*[process_image(synth.synth_image(h, w_, 40 + i)) for i, (h, w_) in enumerate([(256, 320), (320, 256), (192, 256)])])
    pv, grid = np.concatenate(pvs, 0), np.concatenate(grids, 0)
    ids = synth.vl_prompt_ids(cfg, grid, 7)
    data = [pv, grid, None, None, None]
    got = m.forward_initial(ids, 0, data)[0, 0]
    want = o.forward_initial(ids.reshape(1, -1), 0, data)[0, 0]
    err = float(np.abs(got - want).max())
    assert err <= 1e-3, err
    n = pv.shape[0] // 4 * cfg["vision_config"]["out_hidden_size"]
    emb = m.debug_read("image_embeds", 0, n)
    want_emb, _ = o.visual.forward(pv, grid)
    assert float(np.abs(emb - want_emb.reshape(-1)).max()) <= 1e-4
    S = len(ids)
    g2 = m.forward_step(np.array([5], np.uint32), S)[0, 0]
    w2 = o.forward_step(np.array([[5]]), S)[0, 0]
    assert float(np.abs(g2 - w2).max()) <= 1e-3
    mine = torch.from_numpy(np.stack([got, g2]))
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    assert all(torch.equal(g, gathered[0]) for g in gathered), "ranks disagree"
    if rank == 0:
        print(f"TP{world} vl ok: 3 images sharded over {world} ranks, max abs logit err {err:.2e}")
    m.close()


def run(preset, rank, world, local, dist, torch, B200Model, synth, Qwen3Model, nccl_unique_id):
    uid = [nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    cfg = synth.get_config("qwen3", preset)
    w = synth.make_weights("qwen3", cfg, 0)
    m = B200Model("qwen3", cfg, w, device=local, max_ctx=256, tp_rank=rank, tp_world=world, tp_unique_id=uid[0])
    o = Qwen3Model(cfg, w)
    ids = synth.synth_text_ids(70, min(cfg["vocab_size"], 1000) - 8, 3)
    S = 64
    got = m.forward_initial(ids[:S], 0)[0, 0]
    want = o.forward_initial(ids[:S].reshape(1, -1), 0)[0, 0]
    errs = [float(np.abs(got - want).max())]
    outs = [got]
    for i in range(6):
        got = m.forward_step(ids[S + i:S + i + 1], S + i)[0, 0]
        want = o.forward_step(ids[S + i:S + i + 1].reshape(1, 1), S + i)[0, 0]
        errs.append(float(np.abs(got - want).max()))
        outs.append(got)
    assert max(errs) <= 1e-3, errs
    mine = torch.from_numpy(np.stack(outs))
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    for g in gathered:
        assert torch.equal(g, gathered[0]), "ranks disagree"
    # greedy decode through the fused kernel: partial sums and argmax candidates cross NVLink as tagged packets; the
    # tokens must equal the oracle's greedy continuation and agree on every rank
    toks = m.decode_steps(5, S + 6, 8)
    assert m.stats()["kernels_per_decode_step"] == 1, "tensor-parallel decode did not take the fused kernel"
    want_toks, tok = [], 5
    for i in range(8):
        lo = o.forward_step(np.array([[tok]]), S + 6 + i)[0, 0]
        tok = int(np.argmax(lo))
        want_toks.append(tok)
    assert list(toks) == want_toks, (toks, want_toks)
    tl = [None] * world
    dist.all_gather_object(tl, toks)
    assert all(t == tl[0] for t in tl)
    if rank == 0:
        print(f"TP{world} {preset} ok: max abs logit err {max(errs):.2e}, kernels/step {m.stats()['kernels_per_decode_step']}")
    m.close()


if __name__ == "__main__":
    main()
