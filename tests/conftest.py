import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def golden_mesh(name):
    """FaceMesh of a reference test mesh, as exported through the reference's own mesh classes."""
    from oracle import orc
    g = load_golden(f"mesh_{name}.npz")
    return orc.FaceMesh(int(g["n_cells"]), g["face_cell"], g["face_area"], g["face_dist"], g["cell_vol"],
                        g["bface_cell"], g["bface_area"], g["bface_dist"])


@pytest.fixture(scope="session")
def square_nb():
    return golden_mesh("square_nb")


@pytest.fixture(scope="session")
def rectangle():
    return golden_mesh("rectangle")


@pytest.fixture(scope="session")
def step():
    """tests/_data/mesh/step.1 of the reference: 79 672 triangles, the largest config-1 mesh."""
    return golden_mesh("step")


@pytest.fixture(scope="session")
def ctx():
    """A device context; GPU tests fail (not skip) when the CUDA library cannot run."""
    import stormruler_b200 as sb
    c = sb.Context(0)
    yield c
    c.close()


def rhs(n):
    """b[k] = sin(0.37 k): the right-hand side of SURVEY.md 8d config 1."""
    return np.sin(0.37 * np.arange(n))
