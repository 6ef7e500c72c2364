"""dgcnn_b200 -- B200-native (sm_100a) DGCNN hot path: fused graph convolutions and
SortPooling behind the Module API of leftthomas/DGCNN's model.py.  CUDA only."""
from .ops import (ACT_NONE, ACT_TANH, NORM_RW, NORM_SYM, Graph, build_graph, graph_conv_bwd,
                  graph_conv_fwd, graph_ptr, sort_pool_bwd, sort_pool_fwd, stack_bwd, stack_bwd_supported, stack_fwd,
                  stack_fwd_supported)
from .nn import (GCNConv, GraphConvolution, Model, SortAggregation, SortPool,
                 classifier_in_features, fused_enabled, graph_conv_stack, remove_self_loops,
                 set_fused, set_custom_tail, custom_tail_enabled)
from .synth import CONFIGS, GraphBatch, collate, make_batch, make_graphs
from .dp import GradBucket, balanced_shards, shard_bounds, shard_ids
from .optim import FlatAdam
from .trainer import FusedTrainer
from .data import (DeviceDataset, Indegree, ResidentBatch, ResidentLoader, epoch_batches, indegree, load_fold, read_tu_dataset,
                   write_tu_dataset)

__version__ = "0.1.0"
