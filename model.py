"""Drop-in shim: `import model` from the repository root resolves to the B200-native implementation
(cfun_b200.model), so the reference driver heart_main.py (`from config import Config; import model; import utils`,
reference heart_main.py:15-17) runs against these layers unmodified."""
from cfun_b200.model import *  # noqa: F401,F403
from cfun_b200 import model as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
