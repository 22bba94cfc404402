"""Middlebury .flo reader / writer with the reference's names (src/flow_io.py, row f4 of SURVEY.md section 8).

File format: 4-byte tag 'PIEH' (= float32 202021.25, little endian), int32 width, int32 height, then
height x width interleaved float32 (u, v).  Host I/O: nothing here touches the device.
"""
import os
import struct

import numpy as np

TAG_FLOAT = 202021.25
TAG_STRING = b"PIEH"


def read_flow_flo(file_name):
    """src/flow_io.py:29-68 -> float32 [H, W, 2]."""
    if os.path.splitext(file_name)[1] != ".flo":
        raise ValueError("readFlowFile: file_name %s should have extension '.flo'" % file_name)
    with open(file_name, "rb") as fid:
        tag, width, height = struct.unpack("<fii", fid.read(12))
        if tag != TAG_FLOAT:
            raise ValueError("readFlowFile(%s): wrong tag (possibly due to big-endian machine?)" % file_name)
        if not (0 < width < 100000 and 0 < height < 100000):
            raise ValueError("readFlowFile(%s): illegal size %d x %d" % (file_name, width, height))
        flow = np.fromfile(fid, np.float32, count=height * width * 2)
    return flow.reshape(height, width, 2)


def write_flow_flo(flow_data, fname):
    """src/flow_io.py:71-109."""
    if os.path.splitext(fname)[1] != ".flo":
        raise ValueError("writeFlowFile: fname %s should have extension '.flo'" % fname)
    flow_data = np.asarray(flow_data)
    if flow_data.ndim != 3 or flow_data.shape[2] != 2:
        raise ValueError("writeFlowFile: image must have two bands")
    height, width = flow_data.shape[:2]
    with open(fname, "wb") as fid:
        fid.write(TAG_STRING)
        fid.write(struct.pack("<ii", width, height))
        fid.write(np.ascontiguousarray(flow_data, dtype=np.float32).tobytes())


def flow_write(flow_data, file_path):
    """src/flow_io.py:12-26: dispatch on the extension (.flo only, like the reference)."""
    ext = os.path.splitext(file_path)[1]
    if ext == ".flo":
        return write_flow_flo(flow_data, file_path)
    raise ValueError("the file type %s is not supported" % ext)
