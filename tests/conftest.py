import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BLS12_381_R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _unhex(v):
    if v is None:
        return None
    if isinstance(v, list):
        return [_unhex(w) for w in v]
    if isinstance(v, str) and v.startswith("0x"):
        return int(v, 16)
    return v


def load_golden():
    with open(os.path.join(ROOT, "tests", "golden", "golden_v1.json")) as fh:
        raw = json.load(fh)

    def conv(o):
        if isinstance(o, dict):
            return {k: conv(v) for k, v in o.items()}
        if isinstance(o, list):
            return [conv(v) for v in o]
        return _unhex(o)

    return conv(raw)


@pytest.fixture(scope="session")
def golden():
    return load_golden()


@pytest.fixture(scope="session")
def bls_p():
    return BLS12_381_R


# First 2^r-th roots of unity of the BLS12-381 scalar field, r = 0..8, as
# listed in the reference's tests/fixtures.py:20-57 (values are field facts).
ROOTS_OF_UNITY = [
    1,
    52435875175126190479447740508185965837690552500527637822603658699938581184512,
    52435875175126190475982595682112313518914282969839895044333406231173219221505,
    28761180743467419819834788392525162889723178799021384024940474588120723734663,
    38476778329304481878022718993882556548812578500290864179952442003245540347252,
    39328881859443649819318207548060215749094715634259317161033277606721139812495,
    7181556051604179363188280445331338471236451149758288711283449754901695186389,
    12058798319732516928593266977629156816578295917773852992327494850156627156852,
    8031134342720706638121837972897357960137225421159210873251699151356237587899,
]


@pytest.fixture(scope="session")
def roots_of_unity():
    return ROOTS_OF_UNITY
