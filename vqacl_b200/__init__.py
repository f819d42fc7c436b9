"""vqacl_b200 — B200-native (sm_100a) implementation of the VQACL hot path: the VL-T5 train step with the
sample-invariant prototype bank, behind the reference's model API. See DESIGN.md."""
from ._lib import VqaclError
from .config import VLT5Config
from .modeling import VLT5, VLT5VQA, VLSeq2SeqLMOutput
from .optim import FusedAdamW, get_constant_schedule_with_warmup
from .data import BatchPrefetcher
from .pipeline import DeviceCollator, PackedFeatureReader, pack_features

__all__ = ["VqaclError", "VLT5Config", "VLT5", "VLT5VQA", "VLSeq2SeqLMOutput", "FusedAdamW", "BatchPrefetcher",
           "get_constant_schedule_with_warmup", "DeviceCollator", "PackedFeatureReader", "pack_features"]
