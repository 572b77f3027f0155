#!/usr/bin/env python3
"""Writes the committed fixtures under tests/golden/ (run in the build container, where /root/reference and
oracle/_ref exist):
  ref_bvh.json                     structure of the REFERENCE builder's trees (oracle/_ref/ref_bvh_dump)
  ref_CBbunny_thumb64x48.npy       thumbnail of the staff golden image reference_results/sky/CBbunny.png
  oracle_cbspheres_48x36_s4_d4.npy oracle regression image (pins determinism of the oracle itself)
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cuda-raytracer_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
REF = "/root/reference"


def main():
    out = {}
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_bvh_dump")
    for name, ml in (("CBbunny", 32), ("CBbunny", 4), ("CBcoil", 32), ("CBgems", 4), ("plane1024", 32)):
        with tempfile.NamedTemporaryFile(suffix=".txt") as f:
            subprocess.check_call([exe, os.path.join(ROOT, "scenes", name + ".b2s"), str(ml), f.name])
            P, N, L = [], [], []
            for line in open(f.name):
                t = line.split()
                if t[0] == "P": P.append(int(t[1]))
                elif t[0] == "N": N.append((int(t[1]), int(t[2])))
                elif t[0] == "L": L.append(int(t[1]))
        P = np.array(P, np.uint64)
        out[f"{name}:{ml}"] = dict(binary_nodes=len(N), wide_levels=L,
                                   order_checksum=int(np.sum(P * (np.arange(len(P), dtype=np.uint64) % 65521)) % (1 << 61)))
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "ref_bvh.json"), "w"), indent=1)

    from PIL import Image
    im = Image.open(os.path.join(REF, "media/pathtracer/reference_results/sky/CBbunny.png")).convert("RGB").resize((64, 48), Image.BOX)
    np.save(os.path.join(ROOT, "tests", "golden", "ref_CBbunny_thumb64x48.npy"), np.asarray(im, np.uint8))

    import orc
    from b2rt._abi import Config
    from b2rt.scene import Scene, place_camera
    sc = Scene.load(os.path.join(ROOT, "scenes", "CBspheres_lambertian.b2s"))
    cam = place_camera(sc, 48, 36)
    img = orc.OracleScene(sc, 4).render(cam, Config(ns_aa=4, max_ray_depth=4, ns_area_light=1, seed=1), 48, 36)
    np.save(os.path.join(ROOT, "tests", "golden", "oracle_cbspheres_48x36_s4_d4.npy"), img)
    print("fixtures written")


if __name__ == "__main__":
    main()
