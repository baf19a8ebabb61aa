"""Drop-in for the hot-path keys of config/config.py:4-35,93,157.  ``ml_collections`` is used when it
is importable; otherwise a minimal ConfigDict (attribute + item access, ``**cfg``) stands in."""

try:  # pragma: no cover
    from ml_collections import ConfigDict
except Exception:  # noqa: BLE001

    class ConfigDict(dict):
        def __getattr__(self, k):
            try:
                return self[k]
            except KeyError as e:
                raise AttributeError(k) from e

        def __setattr__(self, k, v):
            self[k] = v


def load_config(feature_extractor: str = "fbank"):
    cfg = ConfigDict()
    cfg.task = "run"
    cfg.function = "predict"
    cfg.seed = 42
    cfg.device = "gpu"
    cfg.num_devices = 1
    cfg.distributed_training = False
    cfg.feature_extractor = feature_extractor          # [sincnet, fbank, wav2vec2, hubert_base_robust_mgr]
    cfg.frame_shift = 0.01 if cfg.feature_extractor == "fbank" else 0.02
    cfg.supported_models = ["PyanNet", "PyanNet2"]
    cfg.model_name = "PyanNet" if cfg.feature_extractor == "sincnet" else "PyanNet2"
    cfg.model_dict = ConfigDict()
    if cfg.feature_extractor == "sincnet":
        cfg.model_dict.encoding_dim = 60
    elif cfg.feature_extractor == "fbank":
        cfg.model_dict.encoding_dim = 80
    else:
        cfg.model_dict.encoding_dim = 768
    cfg.max_epochs = 20                                 # training keys kept for completeness of the drop-in (config.py:36-40)
    cfg.check_val_every_n_epoch = 3
    cfg.learning_rate = 1e-3
    cfg.batch_size = 80
    cfg.max_duration = 400
    cfg.load_checkpoint = False
    cfg.checkpoint_path = ""
    # synthetic input of BASELINE config 1 (no corpora in this image)
    cfg.clip_seconds = 60.0
    cfg.window_seconds = 5.0
    return cfg
