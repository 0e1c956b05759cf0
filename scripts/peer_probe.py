"""Developer tool (torchrun, >= 2 GPUs): completion timeline of the peer-memory pulls of the C4 sequence-sharded shape.
Prints, per head group, when its blocks have landed relative to the barrier, for several stream counts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from quantumattention_b200 import _native, parallel
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
B, H, S, D = 1, 24, 75600 // world, 128
chunks = parallel.head_chunks(B, H, S)
for ns in (1, 2, 4, 8):
    os.environ["QA_PEER_STREAMS"] = str(ns)
    parallel.PeerGather._cache.clear()
    comm = parallel.PeerGather.get(None, dev, (B, H, S, D), (B, H, S, D), 2)
    main = torch.cuda.current_stream()
    for rep in range(3):
        dist.barrier(); torch.cuda.synchronize()
        t0 = torch.cuda.Event(enable_timing=True); t0.record(main)
        evs = comm.pull(chunks)
        # timing events: one per stream per chunk are plain events; add timed ones by recording after a wait on main
        marks = []
        for group in evs:
            for ev in group:
                main.wait_event(ev)
            m = torch.cuda.Event(enable_timing=True); m.record(main); marks.append(m)
        torch.cuda.synchronize()
    if rank == 0:
        print(f"world {world} streams {ns}: groups landed at (ms after the call): " + " ".join(f"{t0.elapsed_time(m):.3f}" for m in marks), flush=True)
dist.destroy_process_group()
