import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from gr_baz_b200 import synth
from gr_baz_b200.music_doa import music_doa
cfg, W = bench.workload(2)
resp, _ = bench.table_for(cfg)
dev = torch.device("cuda:0")
d_in = synth.gen_windows_torch(cfg, synth.BASE_SEED + 2, 0, W, dev)
def run(env, x, reps=3):
    os.environ["MUSIC_B200_FUSED"] = env
    blk = music_doa(cfg["m"], cfg["n"], cfg["nsamples"], resp, cfg["resolution"])
    outs = []
    for _ in range(reps):
        a = torch.full((W, 1), -7.0, dtype=torch.float32, device=dev)
        l = torch.full((W, 1), -7.0, dtype=torch.float32, device=dev)
        b = torch.full((W, 1), -7, dtype=torch.int32, device=dev)
        blk.process_device(x.data_ptr(), W, a.data_ptr(), l.data_ptr(), None, b.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        outs.append((b.cpu().numpy()[:, 0], l.cpu().numpy()[:, 0]))
    blk.close()
    return outs
ref = run("0", d_in, 1)[0]
fz = run("1", d_in, 4)
for i, (b, l) in enumerate(fz):
    miss = np.nonzero(b == -7)[0]
    bad = np.nonzero((b != ref[0]) & (b != -7))[0]
    print("run", i, "unwritten:", len(miss), miss[:10], " wrong bins:", len(bad), bad[:10], b[bad[:5]], ref[0][bad[:5]], " level mismatches:", int(np.sum(l != ref[1])))
