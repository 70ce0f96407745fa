from .srdataset import (TactileSRDataset, generate_seqs_sr_records, generate_sr_records, save_sr_dataset)  # noqa: F401
