"""Fixture for the frame-loop glue (anerf_b200/frames.py): the reference's cylinder_to_box_2d / kp_to_valid_rays
(core/utils/skeleton_utils.py:607-694, core/utils/ray_utils.py:83-136) on synthetic cameras.  TEST INFRASTRUCTURE;
build container only.   python oracle/make_golden_frames.py [--check]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from anerf_b200 import frames, synthetic  # noqa: E402
from oracle import ref_import  # noqa: E402

CASES = [dict(H=512, W=512, focal=500., angle=0.0, dist=3.0), dict(H=512, W=512, focal=500., angle=1.3, dist=3.0),
         dict(H=240, W=320, focal=210.5, angle=2.9, dist=2.2), dict(H=96, W=64, focal=50., angle=4.0, dist=1.2),
         dict(H=512, W=512, focal=900., angle=0.5, dist=2.0)]


def reference_boxes():
    ref_import.import_reference()
    from core.utils.ray_utils import kp_to_valid_rays
    pose = synthetic.make_pose(11, 24)
    out = []
    for c in CASES:
        c2w = synthetic.orbit_c2w(c["angle"], c["dist"], centre=pose['kps'][0] * np.array([1., 0., 1.])).astype(np.float32)
        rays, valid, _, bboxes = kp_to_valid_rays(torch.as_tensor(c2w)[None], c["H"], c["W"], c["focal"],
                                                  kps=torch.as_tensor(pose["kps"])[None],
                                                  cylinder_params=torch.as_tensor(pose["cyl"])[None])
        out.append(dict(c2w=c2w, tl=np.asarray(bboxes[0][0]), br=np.asarray(bboxes[0][1]), n_valid=len(valid[0]),
                        valid_head=valid[0][:16].numpy(), valid_tail=valid[0][-16:].numpy(),
                        rays_d_head=rays[0][1][:16].numpy()))
    return pose, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    opt = ap.parse_args()
    pose, ref = reference_boxes()
    save = {"cyl": pose["cyl"]}
    for i, (c, r) in enumerate(zip(CASES, ref)):
        idx, (tl, br) = frames.valid_pixels(pose["cyl"], c["H"], c["W"], c["focal"], r["c2w"])
        ok = np.array_equal(tl, r["tl"]) and np.array_equal(br, r["br"]) and len(idx) == r["n_valid"] and \
            np.array_equal(idx[:16].numpy(), r["valid_head"]) and np.array_equal(idx[-16:].numpy(), r["valid_tail"])
        print(f"[frame {i}] {c} box {r['tl']} {r['br']} valid {r['n_valid']} of {c['H'] * c['W']}: {'ok' if ok else 'MISMATCH'}")
        assert ok
        for k, v in r.items():
            save[f"{i}|{k}"] = np.asarray(v)
    save["cases"] = np.array(repr(CASES))
    if not opt.check:
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "frames_box2d.npz"), **save)


if __name__ == "__main__":
    main()
