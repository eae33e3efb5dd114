#!/usr/bin/env python
"""Wall time of the whole IILE render of cfg1 (400 x 400; IISPTIntegrator::render_normal_2 with the reference's
defaults: 16 indirect tasks, 16 direct passes, 32 x 32 hemispheres) through Scene.iile_render, with the identity network,
a small convolutional one and a random-weight U-Net of IISPTNet's size, next to the reference's own classes chained on ONE host thread (ref_harness
iile_render; the reference spreads tasks over its cores, and adds a stdio round trip to a Python child per hemisphere,
which this comparison leaves out).  python tools/time_iile_render.py [tasks] [passes]"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
from conftest import load_product  # noqa: E402
import oracle  # noqa: E402


def unet(K=64):
    """A U-Net of the size of ml/iispt_net.py's IISPTNet (7 -> 3 channels on 32 x 32, three 2x poolings, K = 64 base
    width, skip connections at 32 / 16 / 8), random weights: stands in for the trained network when timing."""
    import torch
    from torch import nn

    def block(cin, cmid, cout, norm, transposed=False):
        conv = nn.ConvTranspose2d if transposed else nn.Conv2d
        layers = [conv(cin, cmid, 3, padding=1), nn.LeakyReLU(0.2)]
        if norm:
            layers.append(nn.BatchNorm2d(cmid))
        layers += [conv(cmid, cout, 3, padding=1), nn.LeakyReLU(0.2)]
        return layers

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            up = lambda: nn.Upsample(scale_factor=2, mode="bilinear")
            self.e0 = nn.Sequential(*block(7, K, K, False))
            self.e1 = nn.Sequential(nn.MaxPool2d(2), *block(K, 2 * K, 2 * K, True))
            self.e2 = nn.Sequential(nn.MaxPool2d(2), *block(2 * K, 4 * K, 4 * K, True))
            self.e3 = nn.Sequential(nn.MaxPool2d(2), *block(4 * K, 8 * K, 4 * K, True), up())
            self.d0 = nn.Sequential(*block(8 * K, 4 * K, 2 * K, True, True), up())
            self.d1 = nn.Sequential(*block(4 * K, 2 * K, K, True, True), up())
            self.d2 = nn.Sequential(*block(2 * K, K, K, False, True), nn.Conv2d(K, 3, 1), nn.ReLU())

        def forward(self, x):
            x0 = self.e0(x)
            x1 = self.e1(x0)
            x2 = self.e2(x1)
            x3 = self.e3(x2)
            x4 = self.d0(torch.cat((x3, x2), 1))
            x5 = self.d1(torch.cat((x4, x1), 1))
            return self.d2(torch.cat((x5, x0), 1))

    return Net()


def main():
    import torch
    tasks = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    passes = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    iile = load_product()
    iile.set_device(0)
    scene = iile.Scene.from_pbrt(oracle.REF_SCENES / "cfg1.pbrt")
    identity = lambda x: x[:, 0:3]
    torch.manual_seed(0)
    conv = torch.nn.Sequential(torch.nn.Conv2d(7, 32, 3, padding=1), torch.nn.LeakyReLU(0.2),
                               torch.nn.Conv2d(32, 32, 3, padding=1), torch.nn.LeakyReLU(0.2),
                               torch.nn.Conv2d(32, 3, 1), torch.nn.ReLU()).cuda().eval()
    out = {"scene": "cfg1 400x400", "indirect_tasks": tasks, "direct_passes": passes}
    big = unet().cuda().eval()
    for name, net in (("identity", identity), ("conv", conv), ("unet_k64", big)):
        scene.iile_render(net, indirect_tasks=2, direct_samples=1)  # warm-up
        torch.cuda.synchronize()
        t = time.time()
        merged, direct, indirect = scene.iile_render(net, indirect_tasks=tasks, direct_samples=passes)
        torch.cuda.synchronize()
        out[f"gpu_seconds_{name}"] = time.time() - t
    out["merged_mean"] = float(merged.mean())
    if oracle.have_ref():
        flat = "/tmp/_iile_time.flat"
        scene.save_flat(flat)
        _, info = oracle.ref_iile_render(flat, "/tmp", tasks, passes, 32, 100.0, 0)
        out["reference_one_thread_seconds"] = info["indirect_seconds"] + info["direct_seconds"]
        out["reference_camera_samples"] = info["camera_samples"]
        out["speedup_vs_one_reference_thread"] = out["reference_one_thread_seconds"] / out["gpu_seconds_identity"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
