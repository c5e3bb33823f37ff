import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / 'tests' / 'golden'


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


@pytest.fixture(scope='session', autouse=True)
def built_library():
    """The C-ABI library is a build artefact (git-ignored): build it once if it is missing so
    that the CPU suite can check the exported surface on a fresh checkout."""
    lib = ROOT / 'pb_chime5_b200' / 'csrc' / 'libgss.so'
    if not lib.exists() or not (lib.parent / 'libgss_dev.so').exists():
        import subprocess
        subprocess.run(['bash', str(ROOT / 'pb_chime5_b200' / 'csrc' / 'build.sh')], check=True)
    return lib


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')
