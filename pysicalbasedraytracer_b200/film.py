"""Progressive / checkpointable film: the whole state of a long render is the fp32 RGBA *sum*
accumulator (A = samples taken) plus the next sample index, because every (pixel, sample) of the Halton /
Sobol sequences is addressable in O(1) (Sampler::SetSampleNumber, Sampler/Sampler.cpp:35-40).  A 4K / 4096-spp
job can therefore be rendered in slices, previewed, checkpointed and resumed - on one GPU or sharded over
several (shard.py) - and the result is bit-identical to rendering the range in one call (per-pixel sums are
accumulated in sample order).

Device memory is a torch tensor (plumbing); all rendering happens in libpbrgpu.so."""
import numpy as np
import torch

from . import _abi as A


class ProgressiveFilm:
    def __init__(self, scene, width, height, device=0, **render_kwargs):
        self.scene, self.W, self.H = scene, width, height
        self.kw = render_kwargs  # integrator, max_depth, rr_threshold, light_strategy
        self.accum = torch.zeros((height, width, 4), dtype=torch.float32, device=torch.device("cuda", device))
        self.stream = torch.cuda.Stream(self.accum.device)
        self.next_sample = 0

    def render(self, n_samples):
        """Add samples [next_sample, next_sample + n_samples) of every pixel."""
        s0, s1 = self.next_sample, self.next_sample + int(n_samples)
        # The accumulator is produced on torch's current stream (zero fill, checkpoint upload) and the C-ABI treats
        # stream 0 as "the library's private non-blocking stream", so the render runs on a stream of our own that is
        # ordered after torch's current stream, and torch's stream is ordered after the render.
        cur = torch.cuda.current_stream(self.accum.device)
        self.stream.wait_stream(cur)
        self.scene.render_device(self.accum.data_ptr(), stream=self.stream.cuda_stream, region=(0, 0, self.W, self.H), s0=s0, s1=s1,
                                 flags=A.RENDER_ACCUMULATE, **self.kw)
        cur.wait_stream(self.stream)
        self.next_sample = s1
        return self

    def mean(self):
        """Linear per-pixel mean RGB so far (numpy, H x W x 3) - a preview at any time."""
        a = self.accum.cpu().numpy()
        return a[..., :3] / np.maximum(a[..., 3:4], 1.0)

    def state_dict(self):
        return {"accum": self.accum.cpu().numpy(), "next_sample": np.int64(self.next_sample),
                "size": np.array([self.W, self.H], dtype=np.int64)}

    def load_state_dict(self, sd):
        assert tuple(int(v) for v in sd["size"]) == (self.W, self.H)
        self.accum.copy_(torch.from_numpy(np.ascontiguousarray(sd["accum"], dtype=np.float32)))
        self.next_sample = int(sd["next_sample"])
        return self

    def save(self, path):
        np.savez_compressed(path, **self.state_dict())

    def load(self, path):
        return self.load_state_dict(np.load(path))
