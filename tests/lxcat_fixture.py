"""A small synthetic cross-section database in the JSON layout `load_lxcat` reads (the reference ships no LXCat data:
SURVEY §8d config 5).  Two targets, an EFFECTIVE record that must be turned into ELASTIC, a 3-body attachment, a
rescaled excitation, and a target whose density is zero (must be skipped)."""
import json
import os


def records():
    def tab(f, es):
        return [[e, f(e)] for e in es]
    es = [0.0, 0.01, 0.1, 0.5, 1.0, 2.0, 5.0, 10.0, 15.6, 20.0, 30.0, 50.0, 100.0]
    step = lambda e0, s: (lambda e: s * max(e - e0, 0.0) / (max(e - e0, 0.0) + 5.0))
    return [
        {"target": "N2", "kind": "EFFECTIVE", "mass_ratio": 1.95e-5, "comment": "momentum transfer", "product": "N2",
         "data": tab(lambda e: 1.0e-19 + 2e-21 * e, es)},
        {"target": "N2", "kind": "EXCITATION", "threshold": 6.17, "comment": "A3", "product": "N2(A3)", "rescale": 0.8,
         "data": tab(step(6.17, 4e-21), es)},
        {"target": "N2", "kind": "IONIZATION", "threshold": 15.6, "comment": "ionisation", "product": "N2+",
         "data": tab(step(15.6, 3e-20), es)},
        {"target": "O2", "kind": "ELASTIC", "mass_ratio": 1.7e-5, "comment": "elastic", "product": "O2",
         "data": tab(lambda e: 6.0e-20, es)},
        {"target": "O2", "kind": "ATTACHMENT", "threshold": 0.0, "comment": "3-body attachment", "product": "O2-",
         "data": tab(lambda e: 1e-43 / (1.0 + e), es), "weight_scale": 2.0},
        {"target": "O2", "kind": "EXCITATION", "threshold": 0.98, "comment": "a1", "product": "O2(a1)",
         "data": tab(step(0.98, 1e-21), es)},
        {"target": "Ar", "kind": "ELASTIC", "mass_ratio": 1.4e-5, "comment": "not present in air", "product": "Ar",
         "data": tab(lambda e: 1e-20, es)},
    ]


def write(path):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as fd:
        json.dump(records(), fd)
    return path
