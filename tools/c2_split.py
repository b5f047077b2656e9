import sys, torch
sys.path.insert(0, "/root/repo")
import bench
import recommendsystemproject_b200 as tt
dev = torch.device("cuda")
wl = bench.workload("c2")
torch.manual_seed(0)
model = tt.TwoTowerModel(tt.GenericTower(wl["cfg"], "user_tower"), tt.GenericTower(wl["cfg"], "item_tower"), *wl["maps"]).to(dev).train()
batch = bench.tree_to(wl["batch_fn"](100), dev)

def graph_time(fn, reps=30):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): g.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

def user():
    model.zero_grad(set_to_none=False)
    out = model.user_tower(batch["user_tower"], model.user_feature_mapping)
    out.sum().backward()
def item():
    model.zero_grad(set_to_none=False)
    a, hn = model._item_side(batch)
    (a.sum() + hn.sum()).backward()
def enc():
    model.zero_grad(set_to_none=False)
    out = model.user_tower.seq_encoder(batch["user_tower"]["sequence"])
    out.sum().backward()
print("user tower fwd+bwd (graph, warm L2): %.3f ms" % graph_time(user))
print("  of which sequence encoder fwd+bwd: %.3f ms" % graph_time(enc))
print("item side  fwd+bwd (graph, warm L2): %.3f ms" % graph_time(item))
