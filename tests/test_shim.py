"""The import shim (spgan_b200.shim): the reference's own import lines resolve to the CUDA drop-ins.
CPU-only; the second test needs the reference tree (build container) and is skipped elsewhere."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("SPGAN_REFERENCE", "/root/reference")


def _run(code, extra_path=()):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    paths = [ROOT] + list(extra_path)
    return subprocess.run([sys.executable, "-c", "import sys; sys.path[:0] = %r; sys.argv = ['train.py']\n%s" % (paths, code)],
                          capture_output=True, text=True, env=env, cwd="/tmp", timeout=600)


def test_shim_redirects_without_reference_tree():
    r = _run("import spgan_b200, spgan_b200.shim as shim\n"
             "shim.install()\n"
             "from Generation.Generator import Generator, AdaptivePointNorm, EdgeBlock\n"
             "from Generation.Discriminator import Discriminator\n"
             "from Common.gradient_penalty import GradientPenalty\n"
             "assert Generator is spgan_b200.Generator and Discriminator is spgan_b200.Discriminator\n"
             "assert GradientPenalty is spgan_b200.GradientPenalty and EdgeBlock is spgan_b200.EdgeBlock\n"
             "from CD_EMD.emd_ import emd_module\n"                      # GAN_metrics.py:15, loss_utils.py:20
             "assert emd_module.emdModule is spgan_b200.emdModule\n"
             "import sys; assert 'metrics' not in sys.modules\n"          # generic names are never shadowed
             "shim.uninstall()\n"
             "assert 'Generation.Generator' not in sys.modules and 'CD_EMD.emd_.emd_module' not in sys.modules\n"
             "print('ok')")
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "Generation")), reason="reference tree not present")
def test_reference_model_module_imports_on_top_of_the_shim():
    """Generation/model.py (the training driver train.py:16 imports) loads unmodified: its
    `from Generation.Generator import Generator` (model.py:30-31) binds our classes, and its import-time-only
    dependencies missing from this image are stood in for."""
    r = _run("import spgan_b200, spgan_b200.shim as shim\n"
             "shim.install(stub_missing=True)\n"
             "import Generation.model as M\n"
             "assert M.Generator is spgan_b200.Generator and M.Discriminator is spgan_b200.Discriminator\n"
             "from Generation.config import opts\n"
             "G = M.Generator(opts); D = M.Discriminator(opts)\n"
             "from Common.network_utils import requires_grad\n"
             "requires_grad(G, False); requires_grad(D, True)\n"
             "assert not any(p.requires_grad for p in G.parameters()) and all(p.requires_grad for p in D.parameters())\n"
             "print('ok', opts.np, opts.nk)", extra_path=[REF])
    assert r.returncode == 0 and "ok 2048 20" in r.stdout, (r.stdout[-500:], r.stderr[-2000:])
