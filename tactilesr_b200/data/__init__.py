from .srdataset import (DevicePrefetcher, TactileSRDataset, generate_seqs_sr_records, generate_sr_records,  # noqa: F401
                        save_sr_dataset)
