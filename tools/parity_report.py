"""GPU parity report against the committed reference goldens (tests/golden/*.pt).
Prints relative-L2 errors of every compared quantity; used to set / justify test tolerances.

    python tools/parity_report.py [--tf32-backbone]
"""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import interactron_b200 as ib  # noqa: E402
from interactron_b200.synthetic import collate_episodes, synthetic_episode  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tf32-backbone", action="store_true")
    args = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = False
    print("backbone conv precision:", "tf32" if args.tf32_backbone else "fp32")
    for mt in ("interactron_random", "interactron"):
        gold = torch.load(os.path.join(G, f"{mt}_predict.pt"))
        cfg = ib.default_config(mt, weights="synthetic")
        model = ib.build_model(cfg.MODEL).cuda().eval()
        loop = model._get_loop()
        loop.backbone_tf32 = args.tf32_backbone
        for ep, g in gold["episodes"].items():
            data = synthetic_episode(ep)
            t0 = time.time()
            out = loop.adapt_detect(data["frames"].cuda(), data["masks"].cuda(), post_frames=(0,), want_trace=True)
            torch.cuda.synchronize()
            dt = time.time() - t0
            t = out["trace"]
            print(f"== {mt} episode {ep}  ({dt*1e3:.1f} ms eager, launches so far {loop.ops.launch_count()},"
                  f" tf32 gemms {loop.ops.n_tf32}, simt gemms {loop.ops.n_simt})")
            print("   pre logits f0     ", rel(t["pre_logits"][0, 0], g["pre_logits_f0"]))
            print("   pre logits f4     ", rel(t["pre_logits"][0, 4], g["pre_logits_f4"]))
            print("   pre boxes         ", rel(t["pre_boxes"][0], g["pre_boxes"]))
            print("   loss vec          ", rel(t["loss_vec"][0], g["loss_vec"]))
            print("   learned loss      ", rel(out["learned_loss"][0], g["learned_loss"]))
            print("   actions           ", rel(out["actions"][0], g["actions"]))
            names = gold["theta_names"]
            gn = torch.stack([loop.theta_pack.view(t["g"], n)[0].norm() for n in names]).cpu()
            e = ((gn - g["g_norms"]).abs() / g["g_norms"].clamp_min(1e-12))
            print("   g per-tensor norm  max rel", e.max().item(), "median", e.median().item(),
                  "worst:", names[int(e.argmax())])
            worst = (0, "")
            worst_tp = (0, "")
            for n, (gr, tp) in g["small"].items():
                eg = rel(loop.theta_pack.view(t["g"], n)[0], gr)
                et = rel(loop.theta_pack.view(t["theta_prime"], n)[0], tp)
                worst = max(worst, (eg, n))
                worst_tp = max(worst_tp, (et, n))
            print("   g small tensors    worst rel", worst)
            print("   theta' small       worst rel", worst_tp)
            step = torch.stack([(loop.theta_pack.view(t["theta_prime"], n)[0] - loop.theta_pack.view(loop.theta, n)[0]).norm()
                                for n in names]).cpu()
            es = ((step - g["theta_step_norms"]).abs() / g["theta_step_norms"].clamp_min(1e-12))
            print("   |theta'-theta| norm max rel", es.max().item())
            print("   POST logits       ", rel(out["pred_logits"][0], g["pred_logits"][0]))
            print("   POST boxes        ", rel(out["pred_boxes"][0], g["pred_boxes"][0]))
            print("   POST box_features ", rel(out["box_features"][0], g["box_features"][0]))
            d = (out["pred_logits"][0].cpu() - g["pred_logits"][0]).abs()
            print("   POST logits max abs err", d.max().item(), "of max |logit|", g["pred_logits"].abs().max().item())
        if mt == "interactron_random":
            eps = list(gold["episodes"].keys())
            both = model.predict(collate_episodes([synthetic_episode(e) for e in eps]))
            for i, e in enumerate(eps):
                print(f"   batched predict ep{e}: logits", rel(both["pred_logits"][i], gold["episodes"][e]["pred_logits"][0]),
                      "boxes", rel(both["pred_boxes"][i], gold["episodes"][e]["pred_boxes"][0]))
        if mt == "interactron":
            acts = torch.load(os.path.join(G, "interactron_actions.pt"))
            data = synthetic_episode(0)
            for s in range(1, 5):
                d = dict(data)
                d["frames"], d["masks"] = data["frames"][:, :s], data["masks"][:, :s]
                print(f"   get_next_action s={s}: mine {model.get_next_action(d)} ref {acts[s]}")
        del model, loop
        torch.cuda.empty_cache()
    base = torch.load(os.path.join(G, "baselines_predict.pt"))
    m = ib.build_model(ib.default_config("single_frame_baseline", weights="synthetic").MODEL).cuda().eval()
    o = m.predict(synthetic_episode(0, frames=1))
    for k, v in base["detr_ep0_1frame"].items():
        print("   detr", k, rel(o[k], v))
    m = ib.build_model(ib.default_config("multi_frame_baseline", weights="synthetic").MODEL).cuda().eval()
    o = m.predict(synthetic_episode(0))
    for k, v in base["detr_multiframe_ep0"].items():
        print("   detr_multiframe", k, rel(o[k], v))


if __name__ == "__main__":
    main()
