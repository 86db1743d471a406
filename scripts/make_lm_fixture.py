"""Extracts the LM object metadata the synthetic LM-O-shaped generator needs.

Reads DATA files (not source) shipped with the reference:
  /root/reference/data/datasets/lm/models_eval/obj_0000NN_keypoints.ply  (9 keypoints, mm)
  /root/reference/data/datasets/lm/models_eval/models_info.json          (diameter, bbox)
and writes casapose_b200/data/lm_models.json.  Run once in the build container; the
GPU box has no /root/reference and only reads the committed json.
"""
import json
import os

SRC = "/root/reference/data/datasets/lm/models_eval"
DST = os.path.join(os.path.dirname(__file__), "..", "casapose_b200", "data", "lm_models.json")


def read_ply_vertices(path):
    with open(path) as f:
        lines = [l.strip() for l in f]
    n = int([l for l in lines if l.startswith("element vertex")][0].split()[-1])
    start = lines.index("end_header") + 1
    return [[float(t) for t in lines[start + i].split()[:3]] for i in range(n)]


def main():
    info = json.load(open(os.path.join(SRC, "models_info.json")))
    out = {}
    for i in range(1, 16):
        name = "obj_%06d" % i
        kp = read_ply_vertices(os.path.join(SRC, name + "_keypoints.ply"))
        assert len(kp) == 9
        m = info[name]
        out[name] = {
            "keypoints": kp,
            "diameter": m["diameter"],
            "min": [m["min_x"], m["min_y"], m["min_z"]],
            "size": [m["size_x"], m["size_y"], m["size_z"]],
        }
    with open(DST, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", os.path.abspath(DST))


if __name__ == "__main__":
    main()
