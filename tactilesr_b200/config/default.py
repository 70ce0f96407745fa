"""The reference's configuration dictionaries (config/default.py:8-98) with the same keys and values, minus that file's
import-time side effects: no hard-coded ``/code`` root (``$TACTILESR_ROOT``, default the working directory), no nvidia-smi
probe and no ``CUDA_VISIBLE_DEVICES`` write -- the device is chosen by the launcher (``torchrun`` -> LOCAL_RANK)."""
from __future__ import annotations

import os

root_path = os.environ.get("TACTILESR_ROOT", os.getcwd())

common_config = {"root_path": root_path, "random_seed": 42, "deterministic": False, "scale_num": 100}

tPSFNet_config = {
    **common_config,
    "train_batch_size": 256, "test_batch_size": 8, "gama": 1.4, "perception_scale": None, "loss_scale": 1e-1,
    "lr": 1e-4, "lr_scheduler_step_size": 1, "checkpoint_period": 1, "lr_scheduler_gamma": 0.8, "weight_decay": 1e-5,
    "epochs": 51, "sample_cnt": 32,
    "dataset_dir": os.path.join(root_path, "data/rotateDataset"),
    "save_dir": os.path.join(root_path, "pth/tPSFNet_no_aug"),
    "is_aug_data": False,
    "inference_test": False,          # the PNG inference hooks need matplotlib: outside the hot path
}

tactileSR_config = {
    **common_config,
    "train_batch_size": 32, "test_batch_size": 8, "lr": 1e-3, "weight_decay": 1e-2, "lr_scheduler_step_size": 2,
    "lr_scheduler_gamma": 0.8, "checkpoint_period": 1, "HR_scale_num": 10, "sensorMaxVaule_factor": 250, "epochs": 51,
    # as train/tactileSR_train.py:224-227 passes them: warmup_by_epoch is NOT forwarded, so the warm-up runs per iteration
    "warmup_t": 2000, "warmup_by_epoch": True, "warmup_mode": "auto", "warmup_init_lr": 1e-5, "warmup_factor": 1e-4,
    "scale_factor": 10, "seqsCnt": 1, "axisCnt": 3, "patternFeatureExtraLayerCnt": 6, "forceFeatureExtraLayerCnt": 1,
    "inference_test": False,
    "save_dir": os.path.join(root_path, "pth/tactileSR_single"),
    "train_dataset_dir": os.path.join(root_path, "data/SRdataset/SRdataset_train.npy"),
    "test_dataset_dir": os.path.join(root_path, "data/SRdataset/SRdataset_test.npy"),
    "val_dataset_dir": os.path.join(root_path, "data/SRdataset/SRdataset_validation.npy"),
}

tactileSeqs_config = {
    **tactileSR_config,
    "seqsCnt": 7, "axisCnt": 3, "lr": 1e-4, "weight_decay": 1e-2, "epochs": 51,
    "load_checkpoint_dir": os.path.join(root_path, "pth/tactileSR_single/checkpoints/epoch_50.pth"),
    "save_dir": os.path.join(root_path, "pth/tactileSeqs_seq_7"),
    "train_dataset_dir": os.path.join(root_path, "data/SeqsDataset/SRdataset_train_32.npy"),
    "test_dataset_dir": os.path.join(root_path, "data/SeqsDataset/SRdataset_test_32.npy"),
    "val_dataset_dir": os.path.join(root_path, "data/SeqsDataset/SRdataset_validation_32.npy"),
}
