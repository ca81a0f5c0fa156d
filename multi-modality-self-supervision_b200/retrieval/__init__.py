"""Label-conditioned image-report retrieval on the B200 engine: drop-in mirror of
/root/reference/Downstream_task/Retrieval/{retrieval.py, full_dset_retrieval.py} for the scoring path."""
from .full_dset_retrieval import (RetrievalScorer, compute_mrr, compute_ranks, compute_recall_precision, data_processing,  # noqa: F401
                                  evaluate, test)
from .retrieval import CXRBertForRetrieval  # noqa: F401
