"""Host-side mirror of the reference's ``models`` package for the registration hot path.

Same module / class / attribute names as /root/reference/models (BUFFERX.py, patch_embedder.py,
patchnet.py, pose_estimator.py) so that ``from models.BUFFERX import BufferX`` resolves to this
implementation when this directory is put first on ``sys.path``.
"""
