"""Import shim: makes the reference's own import lines resolve to spgan_b200, so that
`train.py` / `Generation/model.py` run unchanged on top of the CUDA path.

    import spgan_b200.shim as shim
    shim.install()                       # before `import Generation.model`
    shim.install(stub_missing=True)      # additionally stand in for import-time-only dependencies

What is redirected (module attribute -> replacement):
    Generation.Generator.{Generator, AdaptivePointNorm, EdgeBlock}      Generation/Generator.py:24-261
    Generation.Discriminator.Discriminator                              Generation/Discriminator.py:48-114
    Common.gradient_penalty.GradientPenalty                             Common/gradient_penalty.py:4-37
    CD_EMD.emd_.emd_module / metrics.emd.emd_module .{emdModule, emdFunction}   metrics/emd/emd_module.py:33-71
    {Generation,Common}.modules.{get_edge_features, edgeConv, knn, get_graph_feature, pairwise_dist,
        get_edge_features_xyz}                                          modules.py:629-796
    Common.ops.{knn, get_graph_feature}, Common.pointnet_util.{square_distance, index_points},
    Common.pointconv_util.{square_distance, index_points, knn_point}    
        (only patched into those modules if they are imported at all: they pull in heavy, unused code)

`stub_missing=True` registers empty stand-ins for modules the reference imports at start-up but never
uses in the default training step and that this image lacks (h5py, tensorboardX, imageio, matplotlib,
the py36 CUDA extension wrappers); it never shadows a module that imports fine.  Not part of any timed path.
"""
import importlib
import importlib.abc
import importlib.machinery
import importlib.util
import sys
import types

_REDIRECTS = {
    "Generation.Generator": ("Generator", "AdaptivePointNorm", "EdgeBlock"),
    "Generation.Discriminator": ("Discriminator",),
    "Common.gradient_penalty": ("GradientPenalty",),
    # `from CD_EMD.emd_ import emd_module` (GAN_metrics.py:15, loss_utils.py:20, data_utils.py:10) and
    # `metrics/emd/emd_module.py`: emdModule / emdFunction over the auction kernel
    "CD_EMD.emd_.emd_module": ("emdModule", "emdFunction"),
}
# redirected only when the parent package really is on sys.path (a generic name like `metrics` must not be shadowed
# by an empty stand-in): metrics/emd/emd_module.py:33-71
_REDIRECT_IF_PARENT = {
    "metrics.emd.emd_module": ("emdModule", "emdFunction"),
}
_GRAPH = ("knn", "get_graph_feature", "pairwise_dist", "get_edge_features_xyz")        # modules.py:629-680, 727-776
_PATCH_IF_IMPORTED = {
    "Generation.modules": ("get_edge_features", "edgeConv") + _GRAPH,
    "Common.modules": ("get_edge_features", "edgeConv") + _GRAPH,
    "Common.ops": ("knn", "get_graph_feature"),                                       # Common/ops.py:129-162
    "Common.pointnet_util": ("square_distance", "index_points"),                      # pointnet_util.py:19-59
    "Common.pointconv_util": ("square_distance", "index_points", "knn_point"),        # pointconv_util.py:107-118
}
# import-time-only dependencies of Generation/model.py (model.py:16,27; H5DataLoader.py:3; visu_utils.py:16-19;
# loss_utils.py:12-13,20-21; data_utils.py:8-11)
_STUB_ROOTS = ("h5py", "tensorboardX", "imageio", "matplotlib", "mpl_toolkits", "open3d", "plyfile", "CD_EMD",
               "pointops", "pointops_cuda", "emd", "emd_cuda", "chamferdistcuda")


class _Anything:
    """Attribute sink for stand-in modules: any attribute is a callable / class that does nothing."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Resolves <root>[.anything] for the roots above to an empty stand-in -- appended to sys.meta_path, so it
    only ever answers for modules nothing else could find."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def _ensure_parent(name):
    """Parent packages: the real ones if importable (the reference tree is on sys.path), else empty namespaces."""
    parts = name.split(".")
    for i in range(1, len(parts)):
        pkg = ".".join(parts[:i])
        if pkg in sys.modules:
            continue
        try:
            importlib.import_module(pkg)
        except Exception:
            m = types.ModuleType(pkg)
            m.__path__ = []
            sys.modules[pkg] = m


def install(stub_missing=False):
    """Idempotent.  Returns the list of module names that were (re)directed."""
    import spgan_b200 as pkg
    done = []
    redirects = dict(_REDIRECTS)
    for modname, names in _REDIRECT_IF_PARENT.items():
        try:
            if importlib.util.find_spec(modname.rsplit(".", 1)[0]) is not None:
                redirects[modname] = names
        except (ImportError, ValueError, AttributeError):
            pass
    for modname, names in redirects.items():
        _ensure_parent(modname)
        m = types.ModuleType(modname)
        m.__doc__ = "spgan_b200 drop-in for %s" % modname
        for n in names:
            setattr(m, n, getattr(pkg, n))
        m.__spgan_b200__ = True
        sys.modules[modname] = m
        parent = sys.modules.get(modname.rsplit(".", 1)[0])
        if parent is not None:
            setattr(parent, modname.rsplit(".", 1)[1], m)
        done.append(modname)
    for modname, names in _PATCH_IF_IMPORTED.items():
        m = sys.modules.get(modname)
        if m is not None:
            for n in names:
                if hasattr(m, n):                    # Common.modules lacks some of Generation.modules' helpers
                    setattr(m, n, getattr(pkg, n))
            done.append(modname)
    if stub_missing and not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())          # last: only consulted when the real import machinery fails
        done.append("stub finder for " + ", ".join(_STUB_ROOTS))
    return done


def uninstall():
    for modname in list(_REDIRECTS) + list(_REDIRECT_IF_PARENT):
        m = sys.modules.get(modname)
        if m is not None and getattr(m, "__spgan_b200__", False):
            del sys.modules[modname]
    sys.meta_path[:] = [f for f in sys.meta_path if not isinstance(f, _StubFinder)]
