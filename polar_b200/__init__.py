"""polar_b200 -- B200 (sm_100a) LLR-domain SC / SCL polar decoder behind the reference's
`PolarCode` surface (tavildar/Polar, PolarC/PolarCode.h). See DESIGN.md / INTEGRATION.md."""
from ._lib import PolarB200Error  # noqa: F401
from .code import HostBuffer, PolarCode, pack_bits, unpack_bits  # noqa: F401

__all__ = ["PolarCode", "HostBuffer", "PolarB200Error", "pack_bits", "unpack_bits"]
