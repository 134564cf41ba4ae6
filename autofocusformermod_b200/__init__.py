"""autofocusformermod_b200 -- B200 (sm_100a) implementation of AutoFocusFormer's CLUSTEN neighbourhood-attention hot path.

Drop-in for ``mask2former.modeling.clusten`` (the four autograd Functions) and for the point utilities the AFF backbone
calls.  All compute goes through libclusten_b200.so (C ABI, include/clusten_b200.h); there is no CPU fallback.
"""
from .ops import (CLUSTENQKFunction, CLUSTENAVFunction, CLUSTENWFFunction, WEIGHTEDGATHERFunction,  # noqa: F401
                  MSDETRPCFunction, inverse_neighbour_list)
from .point_utils import (knn_keops, space_filling_cluster, shepard_decay_weights, upsample_feature_shepard,  # noqa: F401
                          topk_select, mask_select, merge_select)

__all__ = ["CLUSTENQKFunction", "CLUSTENAVFunction", "CLUSTENWFFunction", "WEIGHTEDGATHERFunction", "MSDETRPCFunction",
           "knn_keops", "space_filling_cluster", "shepard_decay_weights", "upsample_feature_shepard",
           "topk_select", "mask_select", "merge_select", "inverse_neighbour_list"]
