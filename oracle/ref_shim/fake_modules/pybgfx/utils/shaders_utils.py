"""pybgfx.utils.shaders_utils stand-in (natrix/core/fluid_simulator.py:7).

The real load_shader() runs bgfx's shaderc on `root_path / name`; here the same file was compiled by g++
(oracle/Makefile).  The source must exist where the caller says it is and be the one the build used."""
from enum import Enum
from pathlib import Path

from pybgfx import ShaderHandle, library


class ShaderType(Enum):
    FRAGMENT = "f"
    VERTEX = "v"
    COMPUTE = "c"


def load_shader(name, shader_type, root_path=None):
    path = Path(root_path or ".") / name
    if not path.is_file():
        raise FileNotFoundError(path)
    info = library().nref_build_info().decode()
    source_root = dict(kv.split("=", 1) for kv in info.split(" ") if "=" in kv).get("source", "")
    if source_root and not str(path.resolve()).startswith(str(Path(source_root).resolve())):
        raise RuntimeError(f"{path} is not under the tree this shim was built from ({source_root})")
    return ShaderHandle(0, name=name, path=path)
