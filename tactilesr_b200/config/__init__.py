from .default import tPSFNet_config, tactileSR_config, tactileSeqs_config  # noqa: F401
