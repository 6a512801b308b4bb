"""Device time of the four training kernels for ONE nc_train_round_batch call on N synthetic groups (2 strands x 2
sequences x 100 events each, r73 template + complement models): the size of one EM wave, so kernel variants can be
compared on identical input.  usage: time_train_round.py N [N ...]   (NC_LIB_PATH selects the library)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nanocall_b200 import api, models as M, synth

def main():
    mdl = {m["name"]: m for m in M.load_builtin_models()}
    tt, tc = mdl["r73.t.006.ont.model"]["table"], mdl["r73.c.p1.006.ont.model"]["table"]
    c = api.Context(0, bp_pool_bytes=1 << 30)
    mt, mc = c.register_model(tt, 0), c.register_model(tc, 1)
    rng = np.random.default_rng(3)
    pool = []
    for k in range(64):
        seqs = []
        for sd, tab in ((0, tt), (1, tc)):
            rd = synth.make_read(rng, tab, 200)
            seqs += [(sd, rd["mean"][:100], rd["stdv"][:100], rd["start"][:100]), (sd, rd["mean"][100:], rd["stdv"][100:], rd["start"][100:])]
        pool.append(dict(seqs=seqs, model_id=(mt, mc), pm=(1, 0, 0, 1, 1, 1), st=(0.1, 0.3, 0.1, 0.3)))
    for n in [int(v) for v in sys.argv[1:]]:
        groups = [pool[k % len(pool)] for k in range(n)]
        c.train_round_batch(groups)
        c.train_stats(reset=True)
        out = c.train_round_batch(groups)
        s = c.train_stats(reset=True)
        chk = float(np.sum([o["pm"][0] + o["st"][0] for o in out[:64]]))
        print(f"groups {n:5d}: emission {s['emission_ms']:8.3f} fwbw {s['fwbw_ms']:8.3f} pm_stats {s['pm_stats_ms']:8.3f} st_stats {s['st_stats_ms']:8.3f} ms"
              f"  events {int(s['events'])} waves {int(s['waves'])}  checksum {chk:.6f}", flush=True)
    c.close()

main()
