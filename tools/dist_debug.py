"""torchrun debug: distributed SpMV (each rank) vs the undivided system computed on each rank's own GPU."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import __graft_entry__ as g
pkg = g.load_package()
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("gloo")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
preset = sys.argv[2] if len(sys.argv) > 2 else "S3-hex"
syn = pkg.Synth(preset, n)
s = syn.stride
rs, _ = syn.row_sizes()
bounds = pkg.partition_rows(rs, world)
idt = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    idt.copy_(torch.frombuffer(bytearray(pkg.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(idt, 0)
asm = pkg.Assembly(device=lr)
asm.dist_init(rank, world, idt.numpy().tobytes(), bounds)
asm.dist_synth_to_device(syn)
info = asm.dist_info()
r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
xg = np.random.default_rng(1).standard_normal(syn.nb * s)
asm.upload_x0(xg[r0 * s:r1 * s])
for variant in (1, 0):
    asm.spmv_resident(reps=1, variant=variant)
    q = asm.download_vector(1)
    full = pkg.Assembly(device=lr)
    syn.to_device(full)
    full.upload_x0(xg)
    full.spmv_resident(reps=1, variant=1)
    qf = full.download_vector(1)[r0 * s:r1 * s]
    full.close()
    bad = np.flatnonzero(~np.isfinite(q) | (np.abs(q - qf) > 1e-11 * np.abs(qf).max()))
    print(f"rank {rank} variant {variant} info {info} rows {r0}..{r1} bad entries {bad.size} first {bad[:6] // s} last {bad[-6:] // s if bad.size else []} nan {np.isnan(q).sum()}", flush=True)
b = asm.download_rhs()
print(f"rank {rank} b finite {np.isfinite(b).all()} |b|max {np.abs(b).max()}", flush=True)
d = asm.inverse_diagonal() if False else None
for variant in (0, 1):
    for nssor in (0, 32):
        asm.set_option("spmv_variant", variant)
        asm.upload_x0(None)
        try:
            print(f"rank {rank} variant {variant} nssor {nssor} pcg", asm.pcg_resident(nssor=nssor), flush=True)
        except Exception as e:
            print(f"rank {rank} variant {variant} nssor {nssor} pcg failed {e}", flush=True)
            for w, nm in ((0, 'x'), (2, 'r'), (3, 'p')):
                v = asm.download_vector(w)
                print(f"rank {rank}   {nm}: nan {np.isnan(v).sum()} first nan idx {np.flatnonzero(np.isnan(v))[:4]} max {np.nanmax(np.abs(v))}", flush=True)
dist.barrier()
