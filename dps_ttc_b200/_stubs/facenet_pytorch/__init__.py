"""Stub of facenet_pytorch (absent offline; pretrained weights need the network).  The reference only
constructs these when guidance images are given and sem_guid_scale != 0 (condition_methods.py:119-141)."""


class _Unavailable:
    def __init__(self, *args, **kwargs):
        raise RuntimeError("facenet_pytorch is not installed: pass an `embedder` to ps_semantic instead")


class MTCNN(_Unavailable):
    pass


class InceptionResnetV1(_Unavailable):
    pass
