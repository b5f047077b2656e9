#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
  1. corpus-sharded top-K (tcgen05 path per shard) + all-gather + merge == unsharded top-K, bit-exact rows
  2. row-sharded embedding lookup (owner = id % W, all-to-all of ids and of gathered rows, CUDA gather kernel on the
     owner) == table[ids]
  3. data-parallel training step: replicas stay bit-identical after two steps
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import recommendsystemproject_b200 as tt
from recommendsystemproject_b200 import dist as tdist, ops, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ok = True

# 1. sharded top-K
gen = torch.Generator(device=dev).manual_seed(5)
Nc, D, K, Bq = 60000, 128, 50, 300
corpus = torch.nn.functional.normalize(torch.randn(Nc, D, device=dev, generator=gen), dim=1)
query = torch.nn.functional.normalize(torch.randn(Bq, D, device=dev, generator=gen), dim=1)
bounds = [Nc * r // world for r in range(world + 1)]
s, i = tdist.sharded_topk(query, corpus[bounds[rank]:bounds[rank + 1]].contiguous(), K, rank, world, bounds[:-1],
                          topk_fn=lambda q, e, k, off: ops.score_topk(q, e, k, off, precision="bf16"))
s_ref, i_ref = ops.score_topk(query, corpus, K)
t1 = bool(torch.equal(i, i_ref))
ok &= t1

# 2. sharded lookup
V, Df = 100003, 64
table = torch.randn(V, Df, device=dev, generator=torch.Generator(device=dev).manual_seed(9))
local_tab = table[rank::world].contiguous()
ids = torch.randint(0, V, (257, 7), device=dev, generator=torch.Generator(device=dev).manual_seed(100 + rank))
got = tdist.sharded_lookup(lambda rows: ops.gather_rows(local_tab, rows.reshape(-1), None, None) if rows.numel() else
                           local_tab.new_empty(0, Df), ids, world)
t2 = bool(torch.equal(got, table[ids]))
ok &= t2

# 3. data-parallel step
cfg = synth.config_c2(dropout_scale=0.0)
torch.manual_seed(0)
model = tt.TwoTowerModel(tt.GenericTower(cfg, "user_tower"), tt.GenericTower(cfg, "item_tower"), *synth.MAPS_C2).to(dev).train()
opt = tt.FusedTwoTowerOptimizer(model, lr=5e-4, max_grad_norm=1.0, table_mode="dense")
def mv(o):
    if isinstance(o, torch.Tensor): return o.to(dev)
    if isinstance(o, dict): return {k: mv(v) for k, v in o.items()}
    return [mv(v) for v in o]
batch = mv(synth.make_batch_c2(128, 20, 3, seed=50 + rank))
step = tdist.DataParallelStep(model, opt, batch, 0.15)
for _ in range(2):
    loss = step()
torch.cuda.synchronize()
p = opt.flat_p.clone()
lo, hi = p.clone(), p.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
t3 = bool(torch.equal(lo, hi)) and bool(torch.isfinite(loss))
ok &= t3
# 4. ShardedEmbeddingBag: exchange="rows": pooled lookup bitwise == unsharded kernel on the full table; exchange="pooled"
#    (owner-side partial sums): equal to rounding, and both exchanges leave the SAME updated shard (to rounding);
#    untouched rows do not move
Vb, Db, Lb, Bb = 200003, 128, 50, 1000
fullw = torch.randn(Vb, Db, device=dev, generator=torch.Generator(device=dev).manual_seed(77))
g4 = torch.Generator(device=dev).manual_seed(400 + rank)
idb = torch.randint(1, Vb, (Bb, Lb), device=dev, generator=g4)
idb[torch.arange(Lb, device=dev)[None, :] >= torch.randint(1, Lb + 1, (Bb, 1), device=dev, generator=g4)] = 0
upb = torch.randn(Bb, Db, device=dev, generator=g4)
ref_pool = ops.gather_rows(fullw, idb, "mean", 0)
all_ids = [torch.empty_like(idb) for _ in range(world)]
dist.all_gather(all_ids, idb)
touched = torch.zeros(Vb, dtype=torch.bool, device=dev)
touched[torch.cat(all_ids).reshape(-1)] = True
touched[0] = False
mine = touched[rank::world]
t4 = True
shards = {}
for exch in ("rows", "pooled"):
    bag = tdist.ShardedEmbeddingBag(Vb, Db, rank, world, "mean", 0, device=dev, full_weight=fullw, exchange=exch)
    bag.zero_grad()
    pooled = bag(idb)
    t4 &= bool(torch.equal(pooled, ref_pool)) if exch == "rows" else bool(torch.allclose(pooled, ref_pool, atol=1e-6, rtol=1e-5))
    (pooled * upb).sum().backward()
    coef = tdist.global_clip_coef([bag.sq_norm], 1.0)
    before = bag.weight.clone()
    bag.step(coef, 1e-2, torch.ones(1, dtype=torch.int64, device=dev))
    t4 &= bool(torch.equal(bag.weight[~mine], before[~mine])) and bool((bag.weight[mine] != before[mine]).any(dim=1).all())
    shards[exch] = (bag.weight.clone(), float(coef))
t4 &= abs(shards["rows"][1] - shards["pooled"][1]) < 1e-6 * max(1.0, abs(shards["rows"][1]))
# first Adam step moves every touched element by ~lr * sign(g): compare where the gradient is not vanishing
diff = (shards["rows"][0] - shards["pooled"][0]).abs()
t4 &= bool((diff > 1e-4).float().mean() < 1e-3)
ok &= t4
# 5. global-batch in-batch softmax: all-gather of item embeddings + gradient return == the fused CE kernel run by
#    one GPU on the whole global batch (unique item ids: no cross-rank collisions)
Bg, Dg, Hg, Tg = 1024, 128, 256, 0.05
gg = torch.Generator(device=dev).manual_seed(909)
Ug = torch.nn.functional.normalize(torch.randn(world * Bg, Dg, device=dev, generator=gg), dim=1)
Ig = torch.nn.functional.normalize(torch.randn(world * Bg, Dg, device=dev, generator=gg), dim=1)
Pg = torch.nn.functional.normalize(torch.randn(Hg, Dg, device=dev, generator=gg), dim=1)
idg = torch.arange(1, world * Bg + 1, device=dev)
slg = slice(rank * Bg, (rank + 1) * Bg)
ul, il, pl = (t.clone().requires_grad_(True) for t in (Ug[slg], Ig[slg], Pg))
lossg = tdist.global_inbatch_ce(ul, il, idg[slg], pl, Tg)
lossg.backward()
ur, ir, pr = (t.clone().requires_grad_(True) for t in (Ug, Ig, Pg))
ref = ops.fused_inbatch_ce(ur, ir, idg, None, pr, Tg)[0]
ref.backward()
tot = lossg.detach().clone()
dist.all_reduce(tot)
t5 = abs(float(tot) / world - float(ref)) < 1e-5
t5 &= bool(torch.allclose(ul.grad / world, ur.grad[slg], atol=1e-7, rtol=1e-4))
t5 &= bool(torch.allclose(il.grad / world, ir.grad[slg], atol=1e-7, rtol=1e-4))
gp = pl.grad.clone()
dist.all_reduce(gp)
t5 &= bool(torch.allclose(gp / world, pr.grad, atol=1e-7, rtol=1e-4))
ok &= t5
# 5b. the same over the tensor-core kernel's RECTANGULAR form with item ids that collide ACROSS ranks: the false-negative
#     mask must cover the gathered global batch (one process on the whole batch is the reference)
idc = torch.randint(1, 300, (world * Bg,), device=dev, generator=torch.Generator(device=dev).manual_seed(31))
ul, il = (t.clone().requires_grad_(True) for t in (Ug[slg], Ig[slg]))
lossc = tdist.global_inbatch_ce(ul, il, idc[slg], None, Tg, precision="bf16")
lossc.backward()
ur, ir = (t.clone().requires_grad_(True) for t in (Ug, Ig))
refc = ops.fused_inbatch_ce(ur, ir, idc, None, None, Tg, precision="bf16")[0]
refc.backward()
totc = lossc.detach().clone()
dist.all_reduce(totc)
t5b = abs(float(totc) / world - float(refc)) < 2e-5
t5b &= float((ul.grad / world - ur.grad[slg]).norm() / ur.grad[slg].norm()) < 1e-4
t5b &= float((il.grad / world - ir.grad[slg]).norm() / ir.grad[slg].norm()) < 1e-4
ok &= t5b
# 5c. the single-pass training form (forward + dU in one walk over the logit tiles) through the same all-gather, declared
#     id range: against the three-pass reference above (bf16 tolerances: the two forms round different quantities to bf16)
ul, il = (t.clone().requires_grad_(True) for t in (Ug[slg], Ig[slg]))
fl = torch.zeros(1, dtype=torch.int32, device=dev)
loss1 = tdist.global_inbatch_ce(ul, il, idc[slg], None, Tg, precision="bf16", nan_flags=fl, id_bits=ops.id_bits_for(300),
                                single_pass=True)
loss1.backward()
tot1 = loss1.detach().clone()
dist.all_reduce(tot1)
t5c = int(fl) == 0 and abs(float(tot1) / world - float(refc)) < 2e-5
t5c &= float((ul.grad / world - ur.grad[slg]).norm() / ur.grad[slg].norm()) < 6e-3
t5c &= float((il.grad / world - ir.grad[slg]).norm() / ir.grad[slg].norm()) < 2e-3
ok &= t5c
# 6. the integrated step: row-sharded tables (one batched exchange) + data-parallel towers with global BatchNorm
#    statistics + global in-batch softmax  ==  ONE process running the unsharded model on the global batch
#    (TwoTowerModel.py:95-140, GenericTower.py:234, training_utils.py:51-56); dropout 0, distinct item ids
Bl = 256
cfg_s = synth.config_c3(v_user=40001, v_item=20001, dim=64, dropout=0.0, shard=True)
cfg_u = synth.config_c3(v_user=40001, v_item=20001, dim=64, dropout=0.0, shard=False)
torch.manual_seed(1234)
ref_model = tt.TwoTowerModel(tt.GenericTower(cfg_u, "user_tower"), tt.GenericTower(cfg_u, "item_tower"), *synth.MAPS_C3).to(dev).train()
full_state = {k: v.clone() for k, v in ref_model.state_dict().items()}
sh_model = tt.TwoTowerModel(tt.GenericTower(cfg_s, "user_tower"), tt.GenericTower(cfg_s, "item_tower"), *synth.MAPS_C3).to(dev).train()
sh_model.load_state_dict(full_state)          # full tables are sliced into the local shards
gbatch = synth.make_batch_c3(B=Bl * world, L=30, v_user=40001, v_item=20001, seed=77, unique_items=True)
def cut(o):
    if isinstance(o, torch.Tensor): return o[rank * Bl:(rank + 1) * Bl].contiguous()
    if isinstance(o, dict): return {k: cut(v) for k, v in o.items()}
    return [cut(v) for v in o]
lbatch = mv(cut(gbatch))
# replicated (small) tables stay dense: their gradients ride the all-reduce of the flat buffer; row-sharded ones are
# always touched-rows-only on their owner
# eps = 1e-3 >> |g| keeps Adam's first step linear in g (rounding-level gradient differences stay rounding-level)
opt_s = tt.FusedTwoTowerOptimizer(sh_model, lr=1e-2, eps=1e-3, max_grad_norm=1.0, table_mode="dense")
step_s = tdist.ShardedTrainStep(sh_model, opt_s, lbatch, 0.05, loss_precision="fp32")
loss_s = step_s().clone()
step_s.check_flags()
opt_u = tt.FusedTwoTowerOptimizer(ref_model, lr=1e-2, eps=1e-3, max_grad_norm=1.0, table_mode="sparse")
gb = mv(gbatch)
opt_u.zero_grad()
u_, i_, _ = ref_model(gb)
loss_u = ref_model.compute_loss(u_, i_, item_ids=gb["item_tower"]["sparse"][:, 0], temperature=0.05)
loss_u.backward()
opt_u.step()
torch.cuda.synchronize()
t6a = abs(float(loss_s) - float(loss_u)) < 2e-5 * max(1.0, abs(float(loss_u)))
t6b = abs(float(opt_s.total_norm) - float(opt_u.total_norm)) < 1e-4 * float(opt_u.total_norm)
t6 = t6a and t6b
new_state = sh_model.state_dict()             # gathers the shards (collective)
worst, n_bad, n_all = 0.0, 0, 0
for k, v in ref_model.state_dict().items():
    if v.dtype.is_floating_point:
        d = (new_state[k].float() - v.float()).abs()
        worst = max(worst, float(d.max()))
        n_bad += int((d > 5e-6).sum())
        n_all += d.numel()
t6c = n_bad == 0
t6 &= t6c
# second step: the updated weights of both runs must give the same loss again
loss_s2 = step_s().clone()
opt_u.zero_grad()
u_, i_, _ = ref_model(gb)
loss_u2 = ref_model.compute_loss(u_, i_, item_ids=gb["item_tower"]["sparse"][:, 0], temperature=0.05)
torch.cuda.synchronize()
t6d = abs(float(loss_s2) - float(loss_u2)) < 1e-4 * max(1.0, abs(float(loss_u2)))
t6 &= t6d
print(f"[rank {rank}] integrated: loss {t6a} ({float(loss_s):.6f}/{float(loss_u):.6f}) norm {t6b} ({float(opt_s.total_norm):.6f}/{float(opt_u.total_norm):.6f}) "
      f"params {t6c} ({n_bad}/{n_all} off, worst {worst:.2e}) loss2 {t6d} ({float(loss_s2):.6f}/{float(loss_u2):.6f})", flush=True)
ok &= t6
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"dist_check world={world}: sharded_topk={t1} sharded_lookup={t2} dp_replicas_identical={t3} sharded_bag={t4} global_inbatch_ce={t5} global_ce_cross_rank_mask_tc={t5b} global_ce_single_pass={t5c} integrated_sharded_step={t6} (loss {float(loss_s):.6f} vs {float(loss_u):.6f}, worst param diff {worst:.2e}) all_ranks_ok={bool(flag.item())}")
rc = 0 if flag.item() else 1
dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(rc)      # captured graphs hold NCCL work: tearing the communicator down at interpreter exit can hang
