from .vad_engine import VadModel
