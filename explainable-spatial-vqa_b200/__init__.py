"""B200-native Program Executor inference path of guoyu-zhang/explainable-spatial-vqa.

Import name: `explainable_spatial_vqa_b200` (the repo-root shim of that name maps it onto this directory,
whose on-disk name carries a hyphen).  Modules:

  inference_transformer_iqap                 drop-in for the reference's IQAP module (VQAModel)
  inference_transformer_full_annotation_new  drop-in for the FA module (MultiModalTransformer, inference cache)
  sharding                                   one-process-per-GPU batch sharding + final NCCL gather
  _native                                    ctypes binding of libb200vqa.so (csrc/, include/b200vqa.h)
"""
from . import _native  # noqa: F401

__version__ = "0.1.0"
