from .modules import (ConfidencePointnet, GeometryTransformer, MultiheadAttention, PositionTransformer,  # noqa: F401
                      RefineTemplate, TransformerDecoderLayer, build_network, refine_modules)
