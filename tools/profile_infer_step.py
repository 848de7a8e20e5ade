"""ncu driver: ONE eager inference step (B = 8, 384x1280, MF_PRECISION) between cudaProfilerStart/Stop, after warm-up:
   MF_CUDA_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
       --log-file gpurun_out/infer_launches.csv python tools/profile_infer_step.py
Summarise with tools/summarize_launches.py."""
import os
import sys

import torch

os.environ["MF_CUDA_GRAPH"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from monoflex_b200 import synthetic as syn                      # noqa: E402
from monoflex_b200.config import default_cfg                    # noqa: E402
from monoflex_b200.model.detector import KeypointDetector       # noqa: E402

H, W, B = 384, 1280, int(os.environ.get("B", "8"))
model = KeypointDetector(default_cfg(width=W, height=H))
model.load_state_dict(syn.make_state_dict(0))
model = model.cuda().eval()
tg = syn.make_targets(B, W // 4, H // 4)
targets = [t.to("cuda") for t in syn.make_param_lists(tg)]
x = syn.make_images(B, H, W).cuda()
with torch.no_grad():
    for _ in range(3):
        model(x, targets)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    model(x, targets)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("profiled one", model.precision, "inference step")
