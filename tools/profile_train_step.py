"""ncu driver: ONE eager training step (B = 8, 384x1280) between cudaProfilerStart/Stop, after two warm-up steps.
   ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/train_launches.csv \
       python tools/profile_train_step.py
Summarise with tools/summarize_launches.py."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monoflex_b200 import synthetic as syn                      # noqa: E402
from monoflex_b200.config import default_cfg                    # noqa: E402
from monoflex_b200.model.detector import KeypointDetector       # noqa: E402
from monoflex_b200.train import Trainer                         # noqa: E402

H, W, B = 384, 1280, int(os.environ.get("B", "8"))
cfg = default_cfg(width=W, height=H)
model = KeypointDetector(cfg)
model.load_state_dict(syn.make_state_dict(0))
model = model.cuda()
tr = Trainer(model, cfg)
tg = [t.to("cuda") for t in syn.make_train_param_lists(syn.make_train_targets(B, empty_image=B))]
x = syn.make_images(B, H, W).cuda()
for _ in range(2):
    tr.step(x, tg, sync_log=False)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(x, tg, sync_log=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one train step")
