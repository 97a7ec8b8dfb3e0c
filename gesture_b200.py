"""Import alias: the package directory name required by the project layout contains hyphens, which
Python cannot import directly.  `import gesture_b200` registers that directory as the package
`gesture_b200` (submodules import normally: `gesture_b200.model_creation`, `gesture_b200.generator` ...)."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                    "speech-driven-gesture-generation-using-transformer-based-denoising-diffusion-probabilistic-models_b200")
_spec = importlib.util.spec_from_file_location("gesture_b200", os.path.join(_DIR, "__init__.py"),
                                               submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gesture_b200"] = _mod
_spec.loader.exec_module(_mod)
