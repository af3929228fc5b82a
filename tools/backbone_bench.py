"""Times the frozen ResNet-50-DC5 trunk (cuDNN) for 40 frames under several settings."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import interactron_b200 as ib  # noqa: E402

m = ib.build_model(ib.default_config("interactron_random", weights="synthetic").MODEL).cuda().eval()
body = m.detector.backbone[0].body
x = torch.randn(40, 3, 300, 300, device="cuda")
ref = None
for cl in (True, False):
    for bench in (False, True):
        for tf32 in (False, True):
            torch.backends.cudnn.benchmark = bench
            torch.backends.cudnn.allow_tf32 = tf32
            xi = x.contiguous(memory_format=torch.channels_last) if cl else x.contiguous()
            with torch.no_grad():
                for _ in range(3):
                    y = body(xi)["0"]
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(5):
                    y = body(xi)["0"]
                e1.record()
                torch.cuda.synchronize()
            if ref is None:
                ref = y.clone()
            err = ((y.double() - ref.double()).norm() / ref.double().norm()).item()
            print(f"channels_last={cl!s:5s} cudnn.benchmark={bench!s:5s} tf32={tf32!s:5s}: "
                  f"{e0.elapsed_time(e1)/5:8.2f} ms / 40 frames   rel diff vs first {err:.2e}", flush=True)
