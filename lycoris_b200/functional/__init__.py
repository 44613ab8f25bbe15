from .general import (  # noqa: F401
    FUNC_LIST,
    apply_dora_scale,
    factorization,
    power2factorization,
    rebuild_tucker,
    tucker_weight,
    tucker_weight_from_conv,
)
