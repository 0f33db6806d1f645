"""xcolumns_b200 -- B200-native drop-in for the prediction-optimisation hot path of xCOLUMNs.

Same public names as the reference for that path:
    predict_using_bc_with_0approx, predict_optimizing_macro_{f1_score,recall,precision}_using_bc,
    predict_optimizing_coverage_using_bc, calculate_confusion_matrix, ConfusionMatrix,
    predict_weighted_per_instance, predict_top_k, find_classifier_using_fw,
    RandomizedWeightedClassifier
All heavy work runs in hand-written sm_100a CUDA kernels behind a C ABI
(include/xcolumns_b200.h); there is no CPU fallback.
"""
from .block_coordinate import (  # noqa: F401
    predict_optimizing_coverage_using_bc,
    predict_optimizing_instance_precision_using_bc,
    predict_optimizing_mixed_instance_precision_and_macro_balanced_accuracy_using_bc,
    predict_optimizing_mixed_instance_precision_and_macro_f1_score_using_bc,
    predict_optimizing_mixed_instance_precision_and_macro_gmean_using_bc,
    predict_optimizing_mixed_instance_precision_and_macro_hmean_using_bc,
    predict_optimizing_mixed_instance_precision_and_macro_jaccard_score_using_bc,
    predict_optimizing_mixed_instance_precision_and_macro_precision_using_bc,
    predict_optimizing_mixed_instance_precision_and_macro_recall_using_bc,
    predict_optimizing_macro_balanced_accuracy_using_bc,
    predict_optimizing_macro_f1_score_using_bc,
    predict_optimizing_macro_gmean_using_bc,
    predict_optimizing_macro_hmean_using_bc,
    predict_optimizing_macro_jaccard_score_using_bc,
    predict_optimizing_macro_precision_using_bc,
    predict_optimizing_macro_recall_using_bc,
    predict_using_bc_with_0approx,
)
from .confusion_matrix import ConfusionMatrix, calculate_confusion_matrix, calculate_fn, calculate_fp, calculate_tp  # noqa: F401
from .frank_wolfe import (  # noqa: F401
    RandomizedWeightedClassifier,
    find_classifier_optimizing_macro_f1_score_using_fw,
    find_classifier_optimizing_macro_precision_using_fw,
    find_classifier_optimizing_macro_recall_using_fw,
    find_classifier_optimizing_mixed_instance_precision_and_macro_f1_score_using_fw,
    find_classifier_optimizing_mixed_instance_precision_and_macro_precision_using_fw,
    find_classifier_optimizing_mixed_instance_precision_and_macro_recall_using_fw,
    find_classifier_optimizing_mixed_macro_recall_and_macro_precision_using_fw,
    find_classifier_using_fw,
    predict_using_randomized_weighted_classifier,
)
from .weighted_prediction import (  # noqa: F401
    predict_log_weighted_per_instance,
    predict_optimizing_instance_precision,
    predict_optimizing_instance_propensity_scored_precision,
    predict_optimizing_macro_balanced_accuracy,
    predict_optimizing_macro_recall,
    predict_power_law_weighted_per_instance,
    predict_top_k,
    predict_weighted_per_instance,
)

__version__ = "0.1.0"
