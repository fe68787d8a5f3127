from oracle.monai08 import (  # noqa: F401
    BlendMode,
    PytorchPadMode,
    ensure_tuple_rep,
    fall_back_tuple,
    look_up_option,
    optional_import,
)
