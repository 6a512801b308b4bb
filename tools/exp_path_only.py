import sys, time, numpy as np, torch
sys.path.insert(0, '.')
from nanocall_b200 import api, synth, models
table = models.builtin_model("r73.t.006.ont.model")["table"]
R, N = 1480, 10000
batch = synth.make_batch_uniform(1, table, R, N)
ctx = api.Context(0); mid = ctx.register_model(table, 0)
dev = torch.device("cuda:0")
d = {k: torch.from_numpy(batch[k]).to(dev) for k in ("mean", "stdv", "start")}
st = torch.zeros(R*N, dtype=torch.int16, device=dev); mv = torch.zeros(R*N, dtype=torch.uint8, device=dev)
torch.cuda.synchronize()
import os
modes = (("path_only", dict()),) if os.environ.get("NC_PATH_ONLY") else (("full", dict(d_states=st.data_ptr(), d_moves=mv.data_ptr())), ("path_only", dict()))
for name, kw in modes:
    for it in range(3):
        ctx.viterbi_device(batch["ev_off"], d["mean"].data_ptr(), d["stdv"].data_ptr(), d["start"].data_ptr(), None, mid, **kw)
        ms = ctx.last_kernel_ms()
    print(name, "kernel ms", ms, "Mev/s", R*N/ms/1e3, flush=True)
    print(ctx.viterbi_stats(), flush=True)
