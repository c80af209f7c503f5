from .eval_retrieval import (  # noqa: F401
    compute_metrics,
    compute_nearest_neighbors,
    compute_pr_at_k,
    construct_embeddings_matrix,
    get_nearest_info,
    print_nearest_info,
    retrieve,
    metrics_from_ranks,
    metrics_from_rank_counts,
    retrieve_metrics,
)
from .accumulator import RetrievalAccumulator, save_predictions  # noqa: F401,E402
