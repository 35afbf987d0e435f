"""Regenerates tests/golden/*.npz from the reference itself (oracle/_ref/libtimg_ref.so, i.e. the
UNMODIFIED timg translation units compiled by oracle/Makefile).  Run in the build container
(where /root/reference exists):   python tests/golden/make_golden.py

Inputs are not stored: tests/cases.py regenerates them deterministically.  Outputs are stored
in full (zip-compressed), keyed by case name.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle  # noqa: E402
import cases  # noqa: E402


def main():
    blocks = {}
    for name, case in cases.block_cases():
        outs = cases.run_block_case(lambda q, u, c: oracle.RefBlockCanvas(q, u, c), case)
        for i, o in enumerate(outs):
            blocks[f"{name}/{i}"] = np.frombuffer(o, np.uint8)
    np.savez_compressed(os.path.join(HERE, "blocks.npz"), **blocks)
    comp = {}
    for name, fb, kw in cases.compose_cases():
        comp[name] = oracle.ref_compose_bg(fb, **kw)
    np.savez_compressed(os.path.join(HERE, "compose.npz"), **comp)
    sc = {}
    for name, img, ow, oh, fmt in cases.scale_cases():
        sc[name] = oracle.ref_scale(img, ow, oh, fmt)
    np.savez_compressed(os.path.join(HERE, "scale.npz"), **sc)
    fit = []
    rng = np.random.default_rng(5)
    for _ in range(400):
        iw, ih = int(rng.integers(1, 5000)), int(rng.integers(1, 5000))
        width, height = int(rng.integers(1, 3000)), int(rng.integers(1, 3000))
        cx, cy = [(1, 2), (2, 2), (9, 18), (1, 1)][int(rng.integers(0, 4))]
        st = float(np.float32([1.0, 0.5, 2.0, 0.1, 7.0, 1.0, 0.8889][int(rng.integers(0, 7))]))
        fl = [int(v) for v in rng.integers(0, 2, 5)]
        args = (iw, ih, width, height, cx, cy, st, *fl)
        r = oracle.calc_fit(iw, ih, width, height, cx, cy, st, *map(bool, fl), impl=oracle.ref().ref_calc_fit)
        fit.append(list(args) + [int(r[0]), r[1], r[2]])
    np.savez_compressed(os.path.join(HERE, "fit.npz"), rows=np.array(fit, np.float64))
    total = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE) if f.endswith(".npz"))
    print(f"wrote {len(blocks)} block outputs, {len(comp)} compose outputs, {len(fit)} fit rows; {total} bytes")


if __name__ == "__main__":
    main()
