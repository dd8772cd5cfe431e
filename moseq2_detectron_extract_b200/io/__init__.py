from .video import RawDepthSession, get_raw_info, read_frames_raw  # noqa: F401
