"""clibd_b200 -- B200-native (sm_100a) implementation of the CLIBD data-parallel hot path.

Host-side mirror of the reference interfaces for that path:

    from clibd_b200 import ContrastiveLoss, ClipLoss            # bioscanclip/model/loss_func.py
    from clibd_b200 import make_prediction, inference_and_print_result, \
        top_k_micro_accuracy, top_k_macro_accuracy, find_closest_match   # bioscanclip/util/util.py
    from clibd_b200 import get_feature_and_label, get_features_and_label  # inference_epoch.py / util.py (device-resident)
    from clibd_b200 import softmax_mean                                  # dna_encoder.py:137 (BarcodeBERT head)
    from clibd_b200 import info_nce_loss, InfoNCELoss                    # bioscanclip/util/simclr.py:64-92,119

Everything computes through the C ABI of include/clibd_b200.h (clibd_b200/lib/libclibd_b200.so,
built by ``python -m clibd_b200._build``); there is no CPU or eager-PyTorch fallback.
"""
from .loss import ClipLoss, ContrastiveLoss, construct_label_metrix, gather_features, pair_weights  # noqa: F401
from .retrieval import (  # noqa: F401
    LEVELS,
    All_TYPE_OF_FEATURES_OF_KEY,
    All_TYPE_OF_FEATURES_OF_QUERY,
    find_closest_match,
    inference_and_print_result,
    knn_search,
    make_prediction,
    top_k_macro_accuracy,
    top_k_micro_accuracy,
)
from .seam import (  # noqa: F401
    EmbeddingStore,
    SoftmaxMeanHead,
    derived_feature_types,
    get_feature_and_label,
    get_features_and_label,
    softmax_mean,
)
from .simclr import InfoNCELoss, info_nce_loss  # noqa: F401

__all__ = [
    "InfoNCELoss", "info_nce_loss",
    "EmbeddingStore", "SoftmaxMeanHead", "derived_feature_types", "get_feature_and_label", "get_features_and_label",
    "softmax_mean",
    "ClipLoss", "ContrastiveLoss", "construct_label_metrix", "gather_features", "pair_weights",
    "LEVELS", "All_TYPE_OF_FEATURES_OF_KEY", "All_TYPE_OF_FEATURES_OF_QUERY", "find_closest_match",
    "inference_and_print_result", "knn_search", "make_prediction", "top_k_macro_accuracy", "top_k_micro_accuracy",
]
