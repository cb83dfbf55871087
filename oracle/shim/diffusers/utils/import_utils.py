import logging as _pylogging
import operator

import torch
from packaging import version


class logging:  # diffusers.utils.logging facade (reference: `from diffusers.utils.import_utils import ..., logging`)
    @staticmethod
    def get_logger(name):
        return _pylogging.getLogger(name)


def is_torch_version(op: str, ver: str) -> bool:
    ops = {">": operator.gt, ">=": operator.ge, "==": operator.eq, "!=": operator.ne, "<=": operator.le, "<": operator.lt}
    return ops[op](version.parse(version.parse(torch.__version__).base_version), version.parse(ver))
