from .fbank import Fbank, FbankConfig
