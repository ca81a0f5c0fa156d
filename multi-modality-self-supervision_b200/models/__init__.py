from .cxrbert_origin import CXRBERT, CXRBertEncoder, ImageBertEmbeddings, ImageTextMatching, BertPreTrainingHeads  # noqa: F401
from .image import ImageEncoder_cnn, ImageEncoder  # noqa: F401
