from .frame import FrameDiffuser  # noqa: F401
from .r3 import R3Diffuser  # noqa: F401
from .so3 import SO3Diffuser  # noqa: F401
