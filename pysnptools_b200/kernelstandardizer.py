"""Kernel standardizers under the reference's module name (``pysnptools/kernelstandardizer/__init__.py:7-100``).

``KernelData.standardize`` takes a :class:`DiagKtoN` (``standardizer/diag_K_to_N.py:54-95``: scale so that the diagonal sums
to ``iid_count``), its trained form, or :class:`Identity`.  They are the classes of :mod:`pysnptools_b200.standardizer`.
"""
from .standardizer import DiagKtoN, DiagKtoNTrained, Identity, Standardizer as KernelStandardizer

__all__ = ["KernelStandardizer", "DiagKtoN", "DiagKtoNTrained", "Identity"]
