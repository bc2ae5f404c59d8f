import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from gr_baz_b200 import synth
from gr_baz_b200.music_doa import music_doa
dev = torch.device("cuda:0")
for cid, W in ((4, 7), (5, 5)):
    cfg = synth.config(cid)
    resp, _ = bench.table_for(cfg)
    d_in = synth.gen_windows_torch(cfg, 123, 0, W, dev)
    Rs = []
    for covn in ("0", "1"):
        os.environ["MUSIC_B200_COVN"] = covn
        blk = music_doa(cfg["m"], cfg["n"], cfg["nsamples"], resp, cfg["resolution"])
        m = cfg["m"]
        a = torch.empty((W, cfg["n"]), dtype=torch.float32, device=dev); l = torch.empty_like(a)
        b = torch.empty((W, cfg["n"]), dtype=torch.int32, device=dev)
        R = torch.zeros((W, m, m, 2), dtype=torch.float64, device=dev)
        try:
            blk.process_device(d_in.data_ptr(), W, a.data_ptr(), l.data_ptr(), None, b.data_ptr(), stream=torch.cuda.current_stream().cuda_stream, d_R=R.data_ptr())
            torch.cuda.synchronize()
        except Exception as e:
            print("config", cid, "covn", covn, "ERROR", e); break
        Rs.append(R.cpu().numpy()); blk.close()
    if len(Rs) == 2:
        d = np.abs(Rs[0] - Rs[1]); ref = np.abs(Rs[0]).max()
        print("config", cid, "max |R_tile - R_covn| / max|R| = %.3e" % (d.max() / ref))
        bad = np.argwhere(d > 1e-9 * ref)
        print("  mismatching entries:", len(bad), bad[:8].tolist())
        if len(bad):
            w = bad[0][0]; print("  R_tile[w][0:2,0:8,0]:", Rs[0][w][0, :8, 0]); print("  R_covn[w][0:2,0:8,0]:", Rs[1][w][0, :8, 0])
