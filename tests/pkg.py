"""Import the product package from its hyphenated directory (stm32h7-yolo_b200/)."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load():
    name = "stm32h7_yolo_b200"
    if name in sys.modules:
        return sys.modules[name]
    path = os.path.join(ROOT, "stm32h7-yolo_b200", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
