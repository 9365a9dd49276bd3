"""Host-buffer pipeline: pinned host clips in, pinned host features out.

The end-to-end form of the batched API for callers whose audio lives in host memory (the decoded
WAV arrays the reference works on): the batch is cut into chunks and each chunk's H2D copy,
kernels and D2H copy are queued on one of a few streams, so transfers of neighbouring chunks
overlap the kernels.  All compute is libb200aa.so; torch provides pinned memory and streams.
"""
import torch

from ._lib import lib, get_plan
from .batch import clip_stats, feature_extraction_batch, NORM_BYTES


class HostPipeline:
    def __init__(self, sampling_rate, window, step, n_samples, max_clips, device=0, chunk_clips=100, n_streams=3,
                 deltas=True, dtype=torch.int16):
        self.fs, self.window, self.step, self.n = int(sampling_rate), int(window), int(step), int(n_samples)
        self.device = torch.device("cuda", device)
        self.deltas = deltas
        self.F = 68 if deltas else 34
        self.T = lib().b200aa_num_frames(self.n, self.window, self.step)
        if self.T <= 0:
            raise ValueError("need at least one array to concatenate")
        self.chunk = max(1, min(int(chunk_clips), int(max_clips)))
        self.max_clips = int(max_clips)
        with torch.cuda.device(self.device):
            self.plan = get_plan(self.fs, self.window, self.step, self.device.index)
            self.streams = [torch.cuda.Stream(self.device) for _ in range(n_streams)]
            self.d_in = [torch.empty((self.chunk, self.n), dtype=dtype, device=self.device) for _ in range(n_streams)]
            self.d_out = [torch.empty((self.chunk, self.F, self.T), dtype=torch.float32, device=self.device) for _ in range(n_streams)]
            self.d_norm = [torch.empty((self.chunk, NORM_BYTES), dtype=torch.uint8, device=self.device) for _ in range(n_streams)]
        self.h_out = torch.empty((self.max_clips, self.F, self.T), dtype=torch.float32).pin_memory()

    def run(self, host_clips, out=None):
        """host_clips: [B, n_samples] CPU tensor (pinned for full-speed copies).  Returns a pinned CPU
        float32 tensor [B, F, T] (a view of an internal buffer unless ``out`` is given)."""
        if isinstance(host_clips, torch.Tensor) is False:
            host_clips = torch.from_numpy(host_clips)
        B = host_clips.shape[0]
        if B > self.max_clips or host_clips.shape[1] != self.n:
            raise ValueError("batch does not match the pipeline's shape")
        h_out = self.h_out if out is None else out
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            s.wait_stream(cur)
        for c, a in enumerate(range(0, B, self.chunk)):
            b = min(B, a + self.chunk)
            n = b - a
            k = c % len(self.streams)
            with torch.cuda.stream(self.streams[k]):
                d_in, d_out = self.d_in[k][:n], self.d_out[k][:n]
                d_in.copy_(host_clips[a:b], non_blocking=True)
                norm = clip_stats(d_in, out=self.d_norm[k][:n])
                feature_extraction_batch(d_in, self.fs, self.window, self.step, deltas=self.deltas, out=d_out,
                                         norm=norm, plan=self.plan)
                h_out[a:b].copy_(d_out, non_blocking=True)
        for s in self.streams:
            s.synchronize()
        return h_out[:B]
