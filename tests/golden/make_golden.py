#!/usr/bin/env python
"""Generate the committed golden fixtures (run in the BUILD container only).

Needs ``/root/reference`` (read-only) and ``oracle/_ref`` (the unmodified
reference header compiled by ``oracle/Makefile``).  Neither exists on the GPU
box, hence the outputs are committed:

* ``kat_elements.json`` -- the Maxima-derived element matrices ``Ke`` (2-D 8x8,
  3-D 24x24) and ``Be`` (2-D 3x8, 3-D 6x24) of the reference's known-answer
  tests, PARSED out of ``/root/reference/tests/test_bri17.cpp`` (``Ke << ...;``
  at :344-356 and :372-531, ``Be << ...;`` at :547-552 and :570-598).
* ``ref_vectors.npz`` -- outputs of the compiled reference header on seeded
  inputs: per-mode ``K^``/``B^`` on full small grids, and whole-grid
  ``f^ = K^ u^`` / ``eps^`` for a few shapes (inputs are regenerated from the
  seed by ``oracle.synthetic_u_hat``; only outputs are stored).

Usage: ``python tests/golden/make_golden.py``
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF_TEST = "/root/reference/tests/test_bri17.cpp"

from oracle import oracle  # noqa: E402


def parse_elements():
    src = open(REF_TEST).read()
    blocks = re.findall(r"\b([KB]e)\s*<<\s*([^;]*);", src)
    assert [b[0] for b in blocks] == ["Ke", "Ke", "Be", "Be"], [b[0] for b in blocks]
    names = ["Ke2", "Ke3", "Be2", "Be3"]
    shapes = [(8, 8), (24, 24), (3, 8), (6, 24)]
    out = {}
    for name, shape, (_, body) in zip(names, shapes, blocks):
        vals = [float(t) for t in body.replace("\n", " ").split(",")]
        assert len(vals) == shape[0] * shape[1], (name, len(vals))
        out[name] = {"shape": list(shape), "values": vals}
    return out


# (dim, shape, spacing, mu, nu, seed)
VECTOR_CASES = [
    (2, (3, 4), (1.1, 1.2), 5.6, 0.3, 11),
    (3, (3, 4, 5), (1.1, 1.2, 1.3), 5.6, 0.3, 12),
    (2, (64, 64), (1.1, 1.2), 5.6, 0.3, 1),          # BASELINE config 1
    (3, (16, 12, 10), (1.1, 1.2, 1.3), 5.6, 0.3, 13),
    (3, (8, 8, 33), (0.125, 0.125, 1.0 / 33), 1.0, 0.3, 14),
]


def main():
    r = oracle.ref()
    if r is None:
        raise SystemExit("oracle/_ref missing: run `make -C oracle` first")
    elements = parse_elements()
    with open(os.path.join(HERE, "kat_elements.json"), "w") as f:
        json.dump({"source": "parsed from /root/reference/tests/test_bri17.cpp "
                             "(:344-356, :372-531, :547-552, :570-598) by make_golden.py",
                   "mu": 5.6, "nu": 0.3,
                   "matrices": elements}, f, indent=0)

    arrays = {}
    meta = []
    for n, (dim, shape, spacing, mu, nu, seed) in enumerate(VECTOR_CASES):
        L = tuple(float(s) * h for s, h in zip(shape, spacing))
        u_hat = oracle.synthetic_u_hat(dim, shape, seed)
        arrays[f"f_hat_{n}"] = r.apply_modal_stiffness(shape, L, mu, nu, u_hat)
        arrays[f"eps_hat_{n}"] = r.apply_strain_displacement(shape, L, u_hat)
        size = int(np.prod(shape))
        if size <= 4096:
            ks = np.stack(np.unravel_index(np.arange(size), shape), axis=1)
            arrays[f"K_{n}"] = np.stack([r.modal_stiffness(shape, L, mu, nu, k) for k in ks])
            arrays[f"B_{n}"] = np.stack([r.modal_strain_displacement(shape, L, k) for k in ks])
        meta.append({"dim": dim, "shape": list(shape), "L": list(L), "mu": mu,
                     "nu": nu, "seed": seed})
    arrays["meta"] = np.array(json.dumps(meta))
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **arrays)
    print("wrote kat_elements.json and ref_vectors.npz:",
          {k: v.shape for k, v in arrays.items() if k != "meta"})


if __name__ == "__main__":
    main()
