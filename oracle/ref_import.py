"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference from /root/reference.

Used here (CPU container) to (a) pin the oracle restatement in `oracle/rover_oracle.py`
bit-for-bit against the reference's own torch code and (b) generate the golden
fixtures under `tests/golden/`.  /root/reference does not exist on the GPU box, so
everything that uses this module is skipped there.  Never imported by the product.

Recipe follows SURVEY.md section 8(c): third-party packages that the reference pulls
in at import time (Isaac Sim, skrl, open3d, ...) are replaced by empty stub packages;
the handful of hard-wired 'cuda:0' defaults are re-pointed at the CPU.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types

REF_ROOT = os.environ.get("ROVER_REFERENCE_ROOT", "/root/reference")
STUB_ROOTS = ("omni", "pxr", "gym", "carb", "open3d", "pymeshlab", "matplotlib", "skrl",
              "hydra", "omegaconf", "rl_games", "importlib_metadata", "scipy_stub_never")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "omniisaacgymenvs"))


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Dummy()

    def __getattr__(self, n):
        return _Dummy()


class _StubModule(types.ModuleType):
    __all__ = []

    def __getattr__(self, n):
        if n.startswith("__"):
            raise AttributeError(n)
        return type(n, (_Dummy,), {})


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in STUB_ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, m):
        pass


_installed = False


def install():
    """Make `omniisaacgymenvs` importable on a CPU-only box without Isaac Sim."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    sys.meta_path.insert(0, _Finder())
    sys.path[:0] = [REF_ROOT, os.path.join(REF_ROOT, "omniisaacgymenvs")]
    _installed = True


def load(device="cpu"):
    """Return a namespace with the reference's hot-path symbols, patched for `device`."""
    import torch
    install()
    from omniisaacgymenvs.tasks.utils.camera import ray_casting, camera, heightmap_distribution
    from omniisaacgymenvs.tasks.utils.rock_detection import rock_detect
    from omniisaacgymenvs.tasks.utils import kinematics
    from omniisaacgymenvs.tasks.utils.math import tensor_quat_to_euler
    from omniisaacgymenvs.tasks import rover

    ns = types.SimpleNamespace()
    if device == "cpu":
        # callers use the default device='cuda:0' (camera.py:110, rock_detect.py:107)
        ray_casting.ray_distance.__defaults__ = ("cpu", torch.float16)

        def quat_cpu(quats, _src=tensor_quat_to_euler.tensor_quat_to_eul):
            # same body as the reference with 'cuda:0' -> quats.device: executed by
            # re-running the reference source text with the literal swapped.
            return _quat_to_eul_on(quats)
        import inspect
        src = inspect.getsource(tensor_quat_to_euler.tensor_quat_to_eul).replace("'cuda:0'", "quats.device")
        g = {"torch": tensor_quat_to_euler.torch}
        exec(src, g)
        _quat_to_eul_on = g["tensor_quat_to_eul"]
        rover.tensor_quat_to_eul = _quat_to_eul_on
        ns.tensor_quat_to_eul = _quat_to_eul_on
        if not getattr(torch.Tensor.cuda, "_rover_patched", False):
            def _ident(self, *a, **k):
                return self
            _ident._rover_patched = True
            torch.Tensor.cuda = _ident
    else:
        ns.tensor_quat_to_eul = tensor_quat_to_euler.tensor_quat_to_eul
    ns.ray_distance = ray_casting.ray_distance
    ns.Camera = camera.Camera
    ns.Heightmap = heightmap_distribution.Heightmap
    ns.Rock_Detection = rock_detect.Rock_Detection
    ns.Ackermann = kinematics.Ackermann
    ns.RoverTask = rover.RoverTask
    ns.Memory = rover.Memory
    ns.rover_module = rover
    return ns
