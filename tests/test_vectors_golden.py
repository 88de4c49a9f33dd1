"""Committed golden vectors (tests/golden/vectors.json, made by tests/golden/make_vectors.py from the oracle):
CPU: the oracle still reproduces them bit for bit; GPU: the CUDA path reproduces them, no oracle in the loop."""
import importlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
import golden_cases as GC

META = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "vectors.json")))


def _case(name):
    gf, extra = GC.inputs(O, name)
    assert GC.input_hash(gf, extra) == META[name]["input_sha256"], \
        "the seeded input generator no longer produces the inputs the golden vectors were made from (numpy RNG stream?)"
    return gf, extra


def _check(name, out):
    gold = META[name]["outputs"]
    assert sorted(out) == sorted(gold)
    for k, v in out.items():
        if GC.digest(v) != gold[k]["sha256"]:
            diff = np.abs(np.asarray(GC.sample(v)) - np.asarray(gold[k]["sample"])).max()
            raise AssertionError("%s/%s differs from the golden vector (sum %r vs %r, max |diff| on the sample block %r)"
                                 % (name, k, float(v.sum()), gold[k]["sum"], diff))


def test_golden_covers_every_case():
    assert sorted(META) == sorted(GC.CASES)


@pytest.mark.parametrize("name", sorted(GC.CASES))
def test_oracle_reproduces_golden(name):
    gf, extra = _case(name)
    _check(name, GC.run_oracle(O, name, gf, extra))


@pytest.mark.gpu
@pytest.mark.parametrize("schedule", [0, 2])
@pytest.mark.parametrize("name", sorted(GC.CASES))
def test_device_reproduces_golden(name, schedule):
    """whole arrays: the result on the interior, halos and level jpk returned as received (exchanged halos for tra_nxt)"""
    N = importlib.import_module("nemo-fmi-devel_b200")
    gf, extra = _case(name)
    _check(name, GC.run_device(N, name, gf, extra, schedule))
