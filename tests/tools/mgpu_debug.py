"""Diagnostics (torchrun, 2 ranks): step a sharded 1-D context and a single-GPU context in lockstep and print where they part."""
import ctypes, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "fit-sne_b200"))
import bench_util, fitsne_b200 as fb
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
def fresh_id():
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = (ctypes.c_char * 128)()
        assert fb.load_library().fitsne_nccl_unique_id(buf) == 0
        idt.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    return idt.cpu().numpy().tobytes()
N = int(os.environ.get("MGPU_N", "20001"))
dims = int(os.environ.get("MGPU_D", "1")); df = float(os.environ.get("MGPU_DF", "0.5"))
row, col, val, labels = bench_util.knn_like_graph(N, 8, seed=3)
Y0 = bench_util.clustered_embedding(labels, dims, 120.0, seed=5)
t = fb.FitSNE(row, col, val, Y0, df=df, device=local, rank=rank, world=world, nccl_id=fresh_id())
s = fb.FitSNE(row, col, val, Y0, df=df, device=local, flags=int(os.environ.get("SINGLE_FLAGS", "0"))) if rank == 0 else None
for it in range(60):
    alpha = 4.0 if it <= 20 else 1.0
    mom = 0.5 if it <= 20 else 0.8
    t.step(exaggeration=alpha, momentum=mom, learning_rate=500.0, max_step_norm=5.0)
    if s is not None:
        s.step(exaggeration=alpha, momentum=mom, learning_rate=500.0, max_step_norm=5.0)
    if it % 10 == 9 or it < 3:
        kl = t.kl(alpha); Y = t.get_Y(); st = t.stats()
        if s is not None:
            kl1 = s.kl(alpha); Y1 = s.get_Y(); st1 = s.stats()
            print("it %2d  kl %.9g / %.9g   Y rel %.3e  nan %d/%d  B %d/%d M %d/%d  bounds [%.5f,%.5f]/[%.5f,%.5f]" % (
                it, kl, kl1, np.linalg.norm(Y - Y1) / np.linalg.norm(Y1), np.isnan(Y).sum(), np.isnan(Y1).sum(), st["n_boxes"], st1["n_boxes"],
                st["fft_side"], st1["fft_side"], st["min_coord"], st["max_coord"], st1["min_coord"], st1["max_coord"]), flush=True)
t.close()
if s is not None: s.close()
if rank == 0: print("DEBUG_DONE", flush=True)
dist.barrier(); dist.destroy_process_group()
