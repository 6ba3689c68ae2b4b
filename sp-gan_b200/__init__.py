"""spgan_b200: the SP-GAN hot path (kNN graph + EdgeConv generator, PointNet critic, WGAN-GP
penalty and training step) on hand-written sm_100a CUDA behind the reference's module API.

There is no CPU or stock-PyTorch fallback: importing the operator layer loads
libspgan_b200.so and fails loudly when it is missing.
"""
from . import ops  # noqa: F401
from .edge import get_edge_features, edgeConv, EdgeBlock  # noqa: F401
from .generator import Generator, AdaptivePointNorm  # noqa: F401
from .discriminator import Discriminator  # noqa: F401
from .gradient_penalty import GradientPenalty  # noqa: F401
from . import pointnet_util  # noqa: F401
from .pointnet_util import (knn, get_graph_feature, pairwise_dist, square_distance, knn_point,  # noqa: F401
                            index_points, get_edge_features_xyz)
from . import emd  # noqa: F401
from .emd import emdModule, emdFunction, emd_approx  # noqa: F401
from .metrics import pairwise_CD, pairwise_EMD, lgan_mmd_cov, one_nn_accuracy  # noqa: F401
from .train_step import WGANGPTrainer, dis_loss_wgan, gen_loss_wgan, requires_grad  # noqa: F401

__all__ = ["ops", "get_edge_features", "edgeConv", "EdgeBlock", "Generator", "AdaptivePointNorm",
           "Discriminator", "GradientPenalty", "WGANGPTrainer", "dis_loss_wgan", "gen_loss_wgan",
           "requires_grad", "pairwise_CD", "pairwise_EMD", "lgan_mmd_cov", "one_nn_accuracy", "pointnet_util", "knn",
           "get_graph_feature", "pairwise_dist", "square_distance", "knn_point", "index_points",
           "get_edge_features_xyz", "emd", "emdModule", "emdFunction", "emd_approx"]
