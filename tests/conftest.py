import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The CPU suite needs liboracle.so and libbrt.so (host-only contexts); build them if absent."""
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    if not os.path.exists(os.path.join(ROOT, "blenderraytracer_b200", "libbrt.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "blenderraytracer_b200", "csrc"), "-s", "-j", "8"])


def load_scene(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def sample_scene():
    return load_scene("sample_scene.json")


@pytest.fixture(scope="session")
def sample_mesh():
    return load_scene("sample_mesh.json")


@pytest.fixture(scope="session")
def kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)

# a checkout of Shinzef/BlenderRayTracer, when this machine has one (the build container; never the GPU box): the tests that execute
# the reference's own JavaScript run against it, everything else uses the committed vectors.  BRT_REFERENCE overrides the location.
REFERENCE = os.environ.get("BRT_REFERENCE", "/root/reference")
REFERENCE_JS = os.path.join(REFERENCE, "js")
HAVE_REFERENCE = os.path.isdir(REFERENCE_JS)

