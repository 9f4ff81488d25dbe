"""The committed golden vectors (tests/golden/oracle_golden.json) are reproduced by the oracle: proof / commitment bytes of NIZK::prove and
SNARK::prove on seeded synthetic instances, MSM results, sumcheck round evaluations.  Guards the checker itself against silent drift."""
import importlib.util
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_reproduces_golden():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    want = json.load(open(os.path.join(HERE, "golden", "oracle_golden.json")))
    got = json.loads(json.dumps(mg.build()))
    assert got == want
    assert want["nizk"][-1]["proof"]["len"] == 9408          # bincode(NIZK) at 2^10
