"""pysnptools_b200: PySnpTools' genotype hot path (decode .bed -> standardize -> kinship) on B200 (sm_100a).

Same names as the reference for this path::

    from pysnptools_b200 import Bed, SnpData, Unit, Beta, SnpKernel
    snpdata = Bed("all.bed", count_A1=False).read(order="F", dtype="float32")
    snpdata = snpdata.standardize(Unit())
    K = SnpKernel(Bed("all.bed", count_A1=False), Unit()).read()
"""
from . import _lib
from .kernelreader import KernelData, SnpKernel
from .snpreader import Bed, SnpData, SnpReader, set_kernel_float64
from .distributedbed import DistributedBed
from .snpmemmap import SnpMemMap
from .standardizer import Beta, BetaTrained, DiagKtoN, Identity, Standardizer, Unit, UnitTrained
from . import kernelreader, kernelstandardizer, snpreader, standardizer, util  # noqa: F401  (the reference's sub-package names)

__all__ = ["Bed", "DistributedBed", "SnpMemMap", "SnpData", "SnpReader", "Unit", "Beta", "UnitTrained", "BetaTrained", "Identity", "DiagKtoN", "Standardizer",
           "SnpKernel", "KernelData", "set_kernel_float64"]
__version__ = "0.1.0"
