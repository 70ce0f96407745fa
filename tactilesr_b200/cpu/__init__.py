"""Host-side training runtime for the hot path: the in-scope pieces of the reference's vendored ``cpu/`` package
(cpu/trainer.py train_one_iter / _log_iter_metrics, cpu/distributed.py), rebuilt for one process per B200."""
from .distributed import (all_gather, gather, get_rank, get_world_size, init_distributed, is_main_process,  # noqa: F401
                          reduce_dict, GradAllReduce)
from .trainer import HookBase, LRWarmupScheduler, MetricStorage, Trainer  # noqa: F401
