from .predict import predict_vad, get_segments, merge_intervals_with_buffer, split_into_windows
