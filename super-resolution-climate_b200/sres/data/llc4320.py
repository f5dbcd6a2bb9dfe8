"""Raw LLC4320 region reader with pinned staging.

Replaces, for the region-of-interest path, the reference's file reader
  SWOTRawDataLoader.load_file / load_region_data / get_dset_time_indices   sres/base/source/swot/raw.py:125-159
  subset_roi                                                                sres/base/source/swot/raw.py:38-45
  mds2d / rearrange                                                         sres/base/source/swot/util.py:3-55
whose per-file work (two ~1 GB np.fromfile reads, a masked scatter, NaN fill, the LLC face unfold and the crop, all
on the host) dominates wall-clock once the training step is fast.  Here the template file (land mask) is read once and
turned into a gather index of the region on the GPU; every data file is then read straight into a pinned staging buffer,
copied to the device in one piece and gathered (byte swap included) by one kernel: (1, ys, xs) fp32 on the device, bit for
bit what the reference's load_file returns.

File naming follows config/dataset/swot_*.yaml: `{dataset_root}/{dataset_files}` with `${dataset.varname}` and
`${dataset.index}` placeholders, template at `{dataset_root}/{template}`.
"""
import ctypes as C
import glob
import os
import re
from typing import Dict, List, Optional

import torch

from sres_b200 import _lib as L


def _expand(pattern: str, varname: str, index) -> str:
    return pattern.replace("${dataset.varname}", str(varname)).replace("${dataset.index}", str(index))


class LLC4320Reader:

    def __init__(self, dataset_root: str, dataset_files: str, template: str, roi: Optional[Dict[str, int]] = None,
                 nx: int = 4320, device: Optional[torch.device] = None):
        self.root, self.files, self.template, self.nx = dataset_root, dataset_files, template, int(nx)
        roi = dict(roi or {})
        self.y0, self.x0 = int(roi.get("y0", 0)), int(roi.get("x0", 0))
        self.ys = int(roi.get("ys", 3 * self.nx - self.y0))
        self.xs = int(roi.get("xs", 4 * self.nx - self.x0))
        self.device = torch.device(device if device is not None else "cuda")
        self.lib = L.lib()
        self.lib.sres_llc_index_workspace_bytes.restype = C.c_size_t
        self._index: Optional[torch.Tensor] = None
        self._staging: Optional[torch.Tensor] = None     # pinned host bytes, reused for every file
        self.n_ocean = -1

    # -- file names -----------------------------------------------------------------------------
    def file_path(self, varname: str, time_index) -> str:
        return os.path.join(self.root, _expand(self.files, varname, time_index))

    def time_indices(self, varname: str) -> List[int]:
        """Indices of the files present for `varname` (raw.py:125-131: glob with '*' in place of the index)."""
        name_rx = re.escape(os.path.basename(self.file_path(varname, "\0"))).replace(re.escape("\0"), r"(\d+)")
        found = []
        for f in glob.glob(self.file_path(varname, "*")):
            m = re.fullmatch(name_rx, os.path.basename(f))
            if m:
                found.append(int(m.group(1)))
        return sorted(found)

    # -- staging ----------------------------------------------------------------------------------
    def _read_pinned(self, path: str) -> torch.Tensor:
        nbytes = os.path.getsize(path)
        if nbytes % 4:
            raise ValueError(f"{path}: size {nbytes} is not a whole number of float32 values")
        if self._staging is None or self._staging.numel() < nbytes:
            self._staging = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        view = self._staging[:nbytes]
        with open(path, "rb", buffering=0) as fh:
            got = fh.readinto(memoryview(view.numpy()))
        if got != nbytes:
            raise IOError(f"{path}: short read ({got} of {nbytes} bytes)")
        return view

    def _to_device(self, path: str) -> torch.Tensor:
        host = self._read_pinned(path)
        dev = torch.empty(host.numel(), dtype=torch.uint8, device=self.device)
        dev.copy_(host, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()   # the staging buffer is reused by the next read
        return dev

    # -- index ------------------------------------------------------------------------------------
    def build_index(self) -> torch.Tensor:
        if self._index is None:
            path = os.path.join(self.root, self.template)
            tmpl = self._to_device(path)
            if tmpl.numel() != 13 * self.nx * self.nx * 4:
                raise ValueError(f"{path}: {tmpl.numel() // 4} values, expected 13*nx^2 = {13 * self.nx * self.nx} for nx = {self.nx}")
            wsb = self.lib.sres_llc_index_workspace_bytes(self.nx)
            with torch.cuda.device(self.device):
                ws = torch.empty(wsb, dtype=torch.uint8, device=self.device)
                index = torch.empty(self.ys * self.xs, dtype=torch.int32, device=self.device)
                n_ocean = torch.zeros(1, dtype=torch.int64, device=self.device)
                L.check(self.lib.sres_llc_build_roi_index(L.ptr(tmpl), self.nx, self.y0, self.ys, self.x0, self.xs, L.ptr(ws),
                                                          C.c_size_t(wsb), L.ptr(index), L.ptr(n_ocean), L.cur_stream()),
                        "sres_llc_build_roi_index")
            self.n_ocean = int(n_ocean.item())
            self._index = index
        return self._index

    # -- data -------------------------------------------------------------------------------------
    def load_file(self, varname: str, time_index: int) -> torch.Tensor:
        """(1, ys, xs) fp32 device tensor = SWOTRawDataLoader.load_file(varname, time_index) (raw.py:133-145)."""
        index = self.build_index()
        path = self.file_path(varname, time_index)
        raw = self._to_device(path)
        if raw.numel() != 4 * self.n_ocean:   # numpy's masked assignment in the reference raises on the same mismatch
            raise ValueError(f"{path}: {raw.numel() // 4} values for {self.n_ocean} ocean points of the template")
        out = torch.empty(1, self.ys, self.xs, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            L.check(self.lib.sres_llc_gather_roi(L.ptr(raw), C.c_int64(self.n_ocean), L.ptr(index), C.c_int64(self.ys * self.xs),
                                                 L.ptr(out), L.cur_stream()), "sres_llc_gather_roi")
        return out

    def load_region_data(self, varnames: List[str], time_index: int) -> torch.Tensor:
        """(C, ys, xs) fp32 on the device: the variables of one time step (raw.py:155-158)."""
        return torch.cat([self.load_file(v, time_index) for v in varnames], dim=0)
