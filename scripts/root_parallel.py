"""Root-parallel MCTS (BASELINE.json config 3, full form): one independent tree per GPU, root child visit/win counts
merged with ONE NCCL all-reduce per move.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/root_parallel.py [iters_per_gpu] [ppl] [lif]
"""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from oponn_b200.engine import RolloutEngine
from oponn_b200.maps import classic42
from oponn_b200.state import ATTACK_START, PHASE_ATTACK, BoardState, Rules
from oponn_b200.tree import SearchTree, TreeOptions, merge_root_children, move_str

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
ppl = int(sys.argv[2]) if len(sys.argv) > 2 else 8
lif = int(sys.argv[3]) if len(sys.argv) > 3 else 16384
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m = classic42(6)
eng = RolloutEngine(m, Rules(), local)
leaf = np.fromfile("tests/golden/classic42_mid_s1.leaf", dtype=np.uint8)
bs = BoardState.unpack(m, leaf); bs.currentPhase, bs.attackState = PHASE_ATTACK, ATTACK_START
for t in range(42):
    if bs.owner[t] == 0: bs.armies[t] = 6
tree = SearchTree(m, Rules(), TreeOptions(agent_id=0, playouts_per_leaf=ppl, leaves_in_flight=lif, seed=1000 + rank), eng)
tree.search(bs, min(iters, 2048))            # warm-up
if world > 1: dist.barrier()
torch.cuda.synchronize(); t0 = time.perf_counter()
ch, st = tree.search(bs, iters)
merged = merge_root_children(ch, torch.device("cuda", local))     # the one exchange of the move
torch.cuda.synchronize(); dt = time.perf_counter() - t0
if world > 1:
    t = torch.tensor([dt], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); dt = float(t.item())
best = tree.choose(merged)
if rank == 0:
    print(json.dumps(dict(n_gpus=world, iterations_per_gpu=iters, playouts_per_leaf=ppl, leaves_in_flight=lif, seconds=dt,
                          playouts_per_move=int(merged["visits"].sum()), playouts_per_s=int(merged["visits"].sum()) / dt,
                          local_playouts=int(ch["visits"].sum()), best=move_str(merged[best]["move"]),
                          agent_win_rate=float(merged["wins"][:, 0].sum() / merged["visits"].sum()),
                          root_children=len(merged))))
tree.close(); eng.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
