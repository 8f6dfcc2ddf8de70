import importlib
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def rtw():
    mod = importlib.import_module("raytracing-one-weekend_b200")
    mod.build()
    return mod


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def port(oracle_mod):
    return oracle_mod.port()


@pytest.fixture(scope="session")
def gpu(rtw):
    if rtw.device_count() < 1:
        pytest.fail("no CUDA device visible to librtw_b200.so: the gpu tests cannot fall back to anything")
    return rtw


GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="session")
def golden():
    import json
    import numpy as np

    def load(name):
        z = np.load(GOLDEN / name, allow_pickle=False)
        d = {k: z[k] for k in z.files if k != "meta"}
        d["meta"] = json.loads(str(z["meta"]))
        return d
    return load


SUZANNE = ROOT / "assets" / "suzanne.obj"
