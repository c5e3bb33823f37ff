"""Host-side mirror of the reference's block interface (pb_chime5/core.py:41-637).

Same class names, constructor fields, call signatures, defaults and error
behaviour as the reference, so that ``pb_chime5/scripts/run.py`` works with only
its import target changed.  Every numeric step is a libgss
CUDA kernel reached through ``pb_chime5_b200.ops``; nothing here computes on the
CPU except the boolean activity framing (a4 in SURVEY.md section 8) and shape
bookkeeping.

Array conventions at this boundary are the reference's:  ``Obs`` (D, T, F)
complex, ``acitivity_freq`` (K, T) bool, masks (K, T, F) / (T, F) float,
``X_hat`` (T, F) complex.  NumPy in -> NumPy out (complex128 / float64 like the
reference); ``torch`` CUDA tensors in -> CUDA tensors out (complex64 / float32,
no host round trip).
"""
from __future__ import annotations

import inspect
import os
from dataclasses import dataclass
from functools import cached_property
from pathlib import Path

import numpy as np
import torch

from . import ops

JSON_PATH = Path(__file__).resolve().parent.parent / 'cache'


# ---------------------------------------------------------------------------
# host <-> device plumbing
# ---------------------------------------------------------------------------

def _device():
    if not torch.cuda.is_available():
        raise RuntimeError('pb_chime5_b200 needs a CUDA device (there is no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device())


def _to_device(x, dtype):
    """NumPy / torch (any device) -> contiguous CUDA tensor of `dtype`; returns (tensor, was_numpy)."""
    if isinstance(x, torch.Tensor):
        return x.to(device=_device(), dtype=dtype).contiguous(), False
    a = np.asarray(x)
    if dtype == torch.complex64:
        a = a.astype(np.complex64, copy=False)
    elif dtype == torch.float32:
        a = a.astype(np.float32, copy=False)
    return torch.from_numpy(np.ascontiguousarray(a)).to(_device()), True


def _from_device(t, was_numpy, np_dtype):
    if not was_numpy:
        return t
    return t.cpu().numpy().astype(np_dtype, copy=False)


def samples_to_stft_frames(samples, size, shift, *, pad=True, fading=False):
    """nara_wpe.utils._samples_to_stft_frames (call site core.py:224-237)."""
    if fading:
        samples = samples + 2 * (size - shift)
    frames = (samples - size + shift) / shift
    return int(np.ceil(frames)) if pad else int(np.floor(frames))


def activity_time_to_frequency(time_activity, stft_window_length, stft_shift, stft_fading, stft_pad=True):
    """Sample-level activity -> frame-level activity ("any sample of the frame").
    Host-side boolean bookkeeping; pb_chime5/database/chime5/database.py:409-472."""
    a = np.asarray(time_activity)
    assert a.dtype != object, (type(time_activity), a.dtype)
    a = a != 0
    if stft_fading:
        p = stft_window_length - stft_shift
        a = np.pad(a, [(0, 0)] * (a.ndim - 1) + [(p, p)], mode='constant')
    n = a.shape[-1]
    if stft_pad:
        extra = (stft_window_length - n) if n < stft_window_length else (-(n - stft_window_length)) % stft_shift
        if extra:
            a = np.pad(a, [(0, 0)] * (a.ndim - 1) + [(0, extra)], mode='constant')
        n = a.shape[-1]
    count = (n - stft_window_length) // stft_shift + 1
    if stft_window_length % stft_shift == 0 and n % stft_shift == 0:
        # a frame is size / shift consecutive hops: any() per hop (one pass over the samples), then the OR
        # of the hops of each frame -- ~4x faster than the cumulative count below on 30 s segments
        hops = a.reshape(a.shape[:-1] + (n // stft_shift, stft_shift)).any(axis=-1)
        r = stft_window_length // stft_shift
        out = hops[..., 0:count].copy()
        for j in range(1, r):
            out |= hops[..., j:j + count]
        return out
    # any() over each frame through a cumulative count (no (T, size) gather)
    c = np.concatenate([np.zeros(a.shape[:-1] + (1,), dtype=np.int32), np.cumsum(a, axis=-1, dtype=np.int32)], axis=-1)
    starts = stft_shift * np.arange(count)
    return (c[..., starts + stft_window_length] - c[..., starts]) > 0


# ---------------------------------------------------------------------------
# blocks
# ---------------------------------------------------------------------------

@dataclass
class WPE:
    """Dereverberation block.  Reference: pb_chime5/core.py:41-88.

    Correlation-build policy (device side, see include/gss.h `gss_wpe_c64_ex`): INT8 tensor cores
    with a float64 re-do of the bins an a-posteriori check flags.  On reverberant, low-noise
    recordings nearly every bin is flagged, so the INT8 attempt is wasted work: the block reads the
    flag statistics of its earlier calls (asynchronously -- it never waits for them) and, once more
    than half of the bins of a call ended on the float64 list, sends the following batches straight
    to the float64 build, probing the INT8 path again every `REPROBE` calls.
    `GSS_WPE_GRAM=f64|i8|i8+redo|auto` in the environment pins the policy."""
    taps: int
    delay: int
    iterations: int
    psd_context: int

    REPROBE = 32            # calls between two INT8 probes while the float64 policy is active

    def _policy(self):
        st = self.__dict__.setdefault('_gram_state', dict(mode=None, calls_in_f64=0, pending=[], last_fraction=None))
        env = os.environ.get('GSS_WPE_GRAM')
        if env:
            return st, (None if env == 'auto' else env), False
        # harvest finished statistics (event.query() does not block)
        keep = []
        for ev, host in st['pending']:
            if ev.query():
                bins, on_list = int(host[0]), int(host[1])
                if bins > 0:
                    st['last_fraction'] = on_list / bins
                    st['mode'] = 'f64' if on_list * 2 > bins else None
                    st['calls_in_f64'] = 0
            else:
                keep.append((ev, host))
        st['pending'] = keep
        if st['mode'] == 'f64':
            st['calls_in_f64'] += 1
            if st['calls_in_f64'] % self.REPROBE == 0:
                return st, None, True                    # probe: INT8 + re-do, with statistics
            return st, 'f64', False
        return st, None, True

    def _run(self, Y, frames=None, info=None, return_f64=False):
        """Y (B,F,D,T) complex64 on the device -> same shape.  frames: valid frames per utterance.
        return_f64: (X complex64, X complex128 before the rounding)."""
        st, mode, want_stats = self._policy()
        stats = torch.zeros(4, dtype=torch.int32, device=Y.device) if want_stats else None
        X = ops.wpe(Y, self.taps, self.delay, self.iterations, self.psd_context, frames=frames,
                    gram_mode=mode, stats=stats, info=info, return_f64=return_f64)
        if stats is not None:
            host = torch.empty(4, dtype=torch.int32, pin_memory=True)
            host.copy_(stats, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            st['pending'] = st['pending'][-3:] + [(ev, host)]
        return X

    @property
    def last_float64_fraction(self):
        """fraction of bins that ended the last harvested call on the float64 list (None: unknown)"""
        return self.__dict__.get('_gram_state', {}).get('last_fraction')

    def __call__(self, Obs, stack=None, debug=False):
        ndim = Obs.ndim
        if ndim == 3:
            assert stack is None, stack
            x, was_np = _to_device(Obs, torch.complex64)                 # (D,T,F)
            out = ops.unpack_fdt_to_dtf(self._run(ops.pack_dtf_to_fdt(x[None])))[0]
        elif ndim == 4:
            x, was_np = _to_device(Obs, torch.complex64)                 # (A,C,T,F)
            A, C, T, F = x.shape
            if stack is True:
                merged = x.reshape(1, A * C, T, F)
                out = ops.unpack_fdt_to_dtf(self._run(ops.pack_dtf_to_fdt(merged)))[0].reshape(A, C, T, F)
            elif stack is False:
                out = ops.unpack_fdt_to_dtf(self._run(ops.pack_dtf_to_fdt(x)))   # arrays are independent batches
            else:
                raise NotImplementedError(stack)
        else:
            raise NotImplementedError(Obs.shape)
        out = _from_device(out, was_np, np.complex128)
        if debug:
            self.locals = dict(Obs=Obs, stack=stack, out=out)
        return out


@dataclass
class Activity:
    """Where the per-session speaker activity comes from (metadata, not on the
    numeric path).  Reference: pb_chime5/core.py:91-141.  The CHiME-5 database
    code is out of scope here and is taken from the reference package when it is
    importable."""
    type: str = 'annotation'
    garbage_class: bool = False
    database_path: str = str(JSON_PATH / 'chime5.json')
    path: str = None

    @cached_property
    def db(self):
        try:
            from pb_chime5.database.chime5 import Chime5
        except ImportError as e:   # pragma: no cover - needs the reference package
            raise RuntimeError(
                'Activity.db needs the CHiME-5 database classes of the reference package '
                '(pb_chime5.database); install fgnt/pb_chime5 next to pb_chime5_b200.') from e
        return Chime5(self.database_path)

    def __getitem__(self, session_id):
        if self.type == 'annotation':
            return _annotation_activity(session_id, self.db, self.garbage_class)
        if self.type == 'path':
            import pickle
            with open(Path(self.path) / f'{session_id}.pkl', 'rb') as fd:
                return pickle.load(fd)
        raise ValueError(self.type)


_ACTIVITY_CACHE = {}


def _annotation_activity(session_id, db, garbage_class):
    key = (session_id, id(db), garbage_class)
    if key not in _ACTIVITY_CACHE:
        from pb_chime5.activity import get_activity   # reference metadata code
        _ACTIVITY_CACHE.clear()                        # keep one session, like lru_cache(1)
        _ACTIVITY_CACHE[key] = get_activity(
            iterator=db.get_datasets(session_id), perspective='array',
            garbage_class=garbage_class, dtype=bool, non_sil_alignment_fn=None,
            debug=False, use_ArrayIntervall=True)[session_id]
    return _ACTIVITY_CACHE[key]


@dataclass
class GSS:
    """Guided source separation: activity-initialised and -constrained CACGMM EM
    per frequency bin.  Reference: pb_chime5/core.py:144-214."""
    iterations: int
    iterations_post: int
    verbose: bool = True

    def _run(self, Y, activity, frames=None, info=None):
        """Y (B,F,D,T) c64, activity (B,K,T_act) uint8/bool -> posterior (B,F,K,T) f32."""
        return ops.cacgmm(Y, activity, self.iterations, self.iterations_post, frames=frames, info=info)

    def __call__(self, Obs, acitivity_freq, debug=False):
        x, was_np = _to_device(Obs, torch.complex64)                     # (D,T,F)
        assert x.ndim == 3, x.shape
        act = torch.as_tensor(np.asarray(acitivity_freq) != 0 if not isinstance(acitivity_freq, torch.Tensor)
                              else acitivity_freq != 0).to(device=x.device, dtype=torch.uint8)
        assert act.ndim == 2, act.shape
        if self.iterations_post == 0:
            # the reference passes an unsupported keyword to CACGMM.predict here (core.py:198-202)
            raise TypeError("predict() got an unexpected keyword argument 'source_activity_mask'")
        Y = ops.pack_dtf_to_fdt(x[None])
        if debug:
            # the reference keeps the fitted models (`learned`, core.py:204-212); here: mixture
            # weights (F,K), class covariances scaled to unit trace (F,K,D,D) and the log
            # determinants the last E-step used (F,K), as host arrays
            post, model = ops.cacgmm(Y, act[None], self.iterations, self.iterations_post, return_model=True)
        else:
            post = self._run(Y, act[None])
        out = ops.unpack_fkt_to_ktf(post)[0]                              # (K,T,F)
        out = _from_device(out, was_np, np.float64)
        if debug:
            learned = {k: v[0].cpu().numpy() for k, v in model.items()}
            self.locals = dict(Obs=Obs, acitivity_freq=acitivity_freq, posterior=out, learned=learned)
        return out


def start_end_context_frames(ex, stft_size, stft_shift, stft_fading):
    """Context length of an example in STFT frames.  Reference: core.py:217-238."""
    start_context_samples = ex['start_orig']['original'] - ex['start']['original']
    end_context_samples = ex['end']['original'] - ex['end_orig']['original']
    assert start_context_samples >= 0, (start_context_samples, ex)
    assert end_context_samples >= 0, (end_context_samples, ex)
    return (samples_to_stft_frames(start_context_samples, stft_size, stft_shift, fading=stft_fading),
            samples_to_stft_frames(end_context_samples, stft_size, stft_shift, fading=stft_fading))


@dataclass
class Beamformer:
    """Mask-based beamformer.  Reference: pb_chime5/core.py:241-278.  In addition
    to the reference's 'mvdrSouden_ban' | 'ch2' | 'sum' the wrapper-level GEV of
    beamforming_wrapper.py:192-208 is selectable as 'gev_ban'."""
    type: str
    postfilter: str

    def _bf_args(self):
        bf = self.type
        if bf in ('mvdrSouden_ban', 'gev_ban', 'mvdrSouden', 'gev'):
            return bf, 0
        if bf == 'ch2':
            return 'ch', 2
        if bf == 'sum':
            return 'sum', 0
        raise NotImplementedError(bf)

    def _dsl(self):
        """Extension: any `get_bf_vector` string of pb_bss (beamformer_wrapper.py:114-227), e.g.
        'rank1_gev+mvdr_souden+ban' or 'wmwf+ban', is accepted as `type` too -> parsed program, else None."""
        if self.type in ('mvdrSouden_ban', 'gev_ban', 'mvdrSouden', 'gev', 'ch2', 'sum'):
            return None
        from .extraction import parse_beamformer
        try:
            return parse_beamformer(self.type)
        except (ValueError, AssertionError):
            raise NotImplementedError(self.type) from None

    def _check_postfilter(self):
        if self.postfilter not in (None, 'mask_mul'):
            raise NotImplementedError(self.postfilter)

    def _run(self, Y, target_mask, distortion_mask):
        """Y (B,F,D,T); masks (B,F,T) f32 -> X_hat (B,F,T) c64."""
        self._check_postfilter()
        bf, arg = self._dsl_type() if self._dsl() is not None else self._bf_args()
        return ops.beamform(Y, target_mask, distortion_mask, bf=bf, postfilter=self.postfilter, bf_arg=arg)

    def _dsl_type(self):
        """`bf_type` word of a DSL program for the C ABI (GSS_BF_PROGRAM, include/gss.h) + N of 'chN'"""
        from . import _lib
        p = self._dsl()
        return (0x100 | _lib.BF_CORES[p['core']] | (_lib.BF_RANK1[p['rank1']] << 4) | (int(p['ban']) << 6)), p['channel']

    def _run_from_posterior(self, Y, posterior, target_index, start_ctx, end_ctx, frames=None, info=None):
        bf, arg = self._dsl_type() if self._dsl() is not None else self._bf_args()
        self._check_postfilter()
        return ops.beamform_from_posterior(Y, posterior, target_index, start_ctx, end_ctx, bf=bf,
                                           postfilter=self.postfilter, bf_arg=arg, frames=frames, info=info)

    def __call__(self, Obs, target_mask, distortion_mask, debug=False):
        if self._dsl() is None:
            self._bf_args()
        self._check_postfilter()
        x, was_np = _to_device(Obs, torch.complex64)
        if x.ndim == 4:                       # '1DTF' (beamforming_wrapper.py:21-22)
            assert x.shape[0] == 1, x.shape
            x = x[0]
        D, T, F = x.shape

        def mask_ft(m):
            m, _ = _to_device(m, torch.float32)
            if m.ndim == 4:
                assert m.shape[0] == 1, m.shape
                m = m[0]
            if m.ndim == 3:                   # (D,T,F): median over the channels (beamforming_wrapper.py:28-30)
                m = m.median(dim=0).values if m.shape[0] % 2 else \
                    0.5 * (m.kthvalue(m.shape[0] // 2, dim=0).values + m.kthvalue(m.shape[0] // 2 + 1, dim=0).values)
            assert m.shape == (T, F), (m.shape, T, F)
            return m.t().contiguous()[None]   # (1,F,T)

        Y = ops.pack_dtf_to_fdt(x[None])
        X = self._run(Y, mask_ft(target_mask), mask_ft(distortion_mask))
        out = ops.unpack_ft_to_tf(X)[0]
        out = _from_device(out, was_np, np.complex128)
        if debug:
            self.locals = dict(Obs=Obs, target_mask=target_mask, distortion_mask=distortion_mask, X_hat=out)
        return out


@dataclass
class Enhancer:
    """STFT -> WPE -> GSS -> context drop -> beamformer -> iSTFT.
    Reference: pb_chime5/core.py:281-571."""
    wpe_block: WPE
    activity: Activity
    gss_block: GSS
    bf_block: Beamformer

    bf_drop_context: bool

    stft_size: int
    stft_shift: int
    stft_fading: bool

    context_samples: int
    multiarray: bool
    reference_array: [None, str]

    @property
    def db(self):
        return self.activity.db

    # ---- transforms (reference layout) -------------------------------------
    def stft(self, x):
        """(..., N) real -> (..., T, F) complex.  core.py:305-312."""
        t, was_np = _to_device(x, torch.float32)
        lead = t.shape[:-1]
        Y = ops.stft(t.reshape(1, -1, t.shape[-1]), self.stft_size, self.stft_shift, self.stft_fading)
        out = Y[0].permute(1, 2, 0).reshape(*lead, Y.shape[3], Y.shape[1])
        return _from_device(out.contiguous(), was_np, np.complex128)

    def istft(self, X):
        """(..., T, F) complex -> (..., samples).  core.py:314-321."""
        t, was_np = _to_device(X, torch.complex64)
        lead = t.shape[:-2]
        Xb = t.reshape(-1, t.shape[-2], t.shape[-1]).permute(0, 2, 1).contiguous()      # (B,F,T)
        out = ops.istft(Xb, self.stft_size, self.stft_shift, self.stft_fading)
        return _from_device(out.reshape(*lead, out.shape[-1]), was_np, np.float64)

    # ---- the device-resident hot path ---------------------------------------
    STAGES = ('wpe', 'cacgmm', 'beamform')

    # Hand-off between WPE and the EM: 'c64' (default; the blocks exchange complex64 tensors in HBM) or
    # 'f64' (the EM sees the dereverberated spectrum unrounded, as in the float64 reference -- removes
    # the one end-to-end difference that is not the rounding of an output, DESIGN.md section 3, at the
    # price of the slower runtime-shape EM kernel).  Not a dataclass field: the constructor mirrors the
    # reference's.  Set `enhancer.handoff = 'f64'` or GSS_HANDOFF=f64.
    handoff = os.environ.get('GSS_HANDOFF', 'c64')

    def enhance_stft_batch(self, Y, activity_freq, target_index, start_ctx, end_ctx, return_masks=False,
                           frames=None, info=None):
        """Y (B,F,D,T) complex64 CUDA (bin-major), activity_freq (B,K,T_act) uint8,
        target_index / start_ctx / end_ctx (B) int32 -> X_hat (B,F,T) complex64
        [, posterior (B,F,K,T) float32].  core.py:524-564 without host round trips.
        frames: optional (B) valid frame counts for a ragged batch padded to T (each utterance
        gives exactly the result it would give alone; padded frames come back as zeros).
        info: optional (3, B) int32 device tensor (`ops.new_info(B, dev, 3)`): the status words of
        the three stages are left there and the call does not synchronise -- the caller checks them
        with `ops.check_info(info, Enhancer.STAGES)` when it fetches the results.  Without it the
        words are checked here (one device synchronisation at the end of the batch)."""
        own = info is None
        if own:
            info = ops.new_info(Y.shape[0], Y.device, stages=3)
        assert self.handoff in ('c64', 'f64'), self.handoff
        Y_em = Y
        if self.wpe_block is not None:
            if self.handoff == 'f64' and self.wpe_block.iterations > 0:
                Y, Y_em = self.wpe_block._run(Y, frames, info=info[0], return_f64=True)
            else:
                Y = Y_em = self.wpe_block._run(Y, frames, info=info[0])
        post = self.gss_block._run(Y_em, activity_freq, frames, info=info[1])
        del Y_em
        if not self.bf_drop_context:
            start_ctx = end_ctx = None
        X = self.bf_block._run_from_posterior(Y, post, target_index, start_ctx, end_ctx, frames, info=info[2])
        if own:
            ops.check_info(info, self.STAGES)
        return (X, post) if return_masks else X

    def enhance_stft_host(self, Obs, acitivity_freq, target_index, start_ctx=None, end_ctx=None,
                          out=None, return_masks=True):
        """Batched hot path on HOST buffers in the reference layout: Obs (B,D,T,F)
        complex64 (pinned torch tensor or NumPy), acitivity_freq (B,K,T_act) bool,
        target_index / start_ctx / end_ctx (B) ints.  Returns host tensors
        X_hat (B,T,F) complex64 and, optionally, masks (B,K,T,F) float32 (context
        frames NOT zeroed, i.e. the posterior as GSS returns it).  Copies in and out
        are issued on the current stream; the call returns after the results landed.
        `out` = optional dict of preallocated pinned host tensors {'X_hat', 'masks'}."""
        dev = _device()
        obs_h = Obs if isinstance(Obs, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(Obs))
        assert obs_h.dtype == torch.complex64 and obs_h.ndim == 4, (obs_h.dtype, obs_h.shape)
        B, D, T, F = obs_h.shape
        x = obs_h.to(dev, non_blocking=True)
        act = torch.as_tensor(acitivity_freq).to(torch.uint8).to(dev, non_blocking=True)
        ivec = lambda v: None if v is None else torch.as_tensor(v, dtype=torch.int32).to(dev, non_blocking=True)
        Y = ops.pack_dtf_to_fdt(x)
        info = ops.new_info(B, dev, stages=3)
        X, post = self.enhance_stft_batch(Y, act, ivec(target_index), ivec(start_ctx), ivec(end_ctx),
                                          return_masks=True, info=info)
        info_h = torch.empty(info.shape, dtype=info.dtype, pin_memory=True)
        info_h.copy_(info, non_blocking=True)
        X_tf = ops.unpack_ft_to_tf(X)
        if out is None:
            out = {}
        xh = out.get('X_hat')
        if xh is None:
            xh = torch.empty(X_tf.shape, dtype=X_tf.dtype, pin_memory=True)
        xh.copy_(X_tf, non_blocking=True)
        res = {'X_hat': xh}
        if return_masks:
            m_ktf = ops.unpack_fkt_to_ktf(post)
            mh = out.get('masks')
            if mh is None:
                mh = torch.empty(m_ktf.shape, dtype=m_ktf.dtype, pin_memory=True)
            mh.copy_(m_ktf, non_blocking=True)
            res['masks'] = mh
        torch.cuda.current_stream().synchronize()
        ops.check_info(info_h, self.STAGES)
        return res

    def enhance_stft_host_stream(self, batches, return_masks=True, reuse_outputs=False):
        """Pipelined variant of `enhance_stft_host` for a sequence of batches: yields one result
        dict per input batch, in order.  `batches` yields tuples
        (Obs (B,D,T,F) complex64 pinned host tensor, acitivity_freq (B,K,T_act), target_index,
        start_ctx, end_ctx).  The host->device copy of batch i+1 and the device->host copy of
        batch i-1 run on side streams while batch i is computed (two input slots in HBM); every
        batch is still copied in and out in full.  Results land in fresh pinned buffers, or, with
        `reuse_outputs=True`, in a ring of three pinned slots per shape (no page-locking inside
        the loop): a yielded result then stays valid until two further results have been yielded."""
        dev = _device()
        compute = torch.cuda.current_stream()
        h2d, d2h = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def upload(item):
            Obs, act, ti, sc, ec = item
            obs_h = Obs if isinstance(Obs, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(Obs))
            assert obs_h.dtype == torch.complex64 and obs_h.ndim == 4, (obs_h.dtype, obs_h.shape)
            ivec = lambda v: None if v is None else torch.as_tensor(v, dtype=torch.int32).to(dev, non_blocking=True)
            with torch.cuda.stream(h2d):
                x = obs_h.to(dev, non_blocking=True)
                a = torch.as_tensor(act).to(torch.uint8).to(dev, non_blocking=True)
                vecs = (ivec(ti), ivec(sc), ivec(ec))
                ready = torch.cuda.Event()
                ready.record(h2d)
            return x, a, vecs, ready

        ring, ring_pos = {}, [0]

        def host_slot(name, like):
            if not reuse_outputs:
                return torch.empty(like.shape, dtype=like.dtype, pin_memory=True)
            key = (name, tuple(like.shape), like.dtype)
            if key not in ring:
                ring[key] = [torch.empty(like.shape, dtype=like.dtype, pin_memory=True) for _ in range(3)]
            return ring[key][ring_pos[0] % 3]

        def download(X_tf, m_ktf, info, done):
            with torch.cuda.stream(d2h):
                d2h.wait_event(done)
                res = {'X_hat': host_slot('X_hat', X_tf), 'info': host_slot('info', info)}
                res['X_hat'].copy_(X_tf, non_blocking=True)
                res['info'].copy_(info, non_blocking=True)
                if m_ktf is not None:
                    res['masks'] = host_slot('masks', m_ktf)
                    res['masks'].copy_(m_ktf, non_blocking=True)
                ring_pos[0] += 1
                fin = torch.cuda.Event()
                fin.record(d2h)
            return res, fin, (X_tf, m_ktf, info)  # keep the device tensors alive until the copy is done

        it = iter(batches)
        try:
            nxt = upload(next(it))
        except StopIteration:
            return
        pending = None
        while nxt is not None:
            x, a, (ti, sc, ec), ready = nxt
            try:
                nxt = upload(next(it))                    # overlaps with the kernels issued below
            except StopIteration:
                nxt = None
            compute.wait_event(ready)
            for t in (x, a, ti, sc, ec):
                if t is not None:
                    t.record_stream(compute)
            Y = ops.pack_dtf_to_fdt(x)
            info = ops.new_info(x.shape[0], dev, stages=3)
            X, post = self.enhance_stft_batch(Y, a, ti, sc, ec, return_masks=True, info=info)
            X_tf = ops.unpack_ft_to_tf(X)
            m_ktf = ops.unpack_fkt_to_ktf(post) if return_masks else None
            done = torch.cuda.Event()
            done.record(compute)
            X_tf.record_stream(d2h)
            info.record_stream(d2h)
            if m_ktf is not None:
                m_ktf.record_stream(d2h)
            cur = download(X_tf, m_ktf, info, done)
            if pending is not None:
                pending[1].synchronize()
                ops.check_info(pending[0]['info'], self.STAGES)    # host words: no device synchronisation
                yield pending[0]
            pending = cur
        pending[1].synchronize()
        ops.check_info(pending[0]['info'], self.STAGES)
        yield pending[0]

    def enhance_observation(self, obs, ex_array_activity, speaker_id, ex=None, debug=False):
        """obs (D, N) samples, ex_array_activity {speaker: (N,) bool} -> x_hat (N',).
        core.py:514-571."""
        t, was_np = _to_device(obs, torch.float32)
        assert t.ndim == 2, t.shape
        Y = ops.stft(t[None], self.stft_size, self.stft_shift, self.stft_fading)        # (1,F,D,T)
        T = Y.shape[3]
        acitivity_freq = activity_time_to_frequency(
            np.array([np.asarray(v) for v in ex_array_activity.values()]),
            stft_window_length=self.stft_size, stft_shift=self.stft_shift,
            stft_fading=self.stft_fading, stft_pad=True)
        act = torch.from_numpy(acitivity_freq.astype(np.uint8))[None].to(Y.device)
        target_speaker_index = tuple(ex_array_activity.keys()).index(speaker_id)
        sc = ec = 0
        if self.bf_drop_context:
            sc, ec = self._context_frames(ex)
        ivec = lambda v: torch.tensor([v], dtype=torch.int32, device=Y.device)
        X, post = self.enhance_stft_batch(Y, act, ivec(target_speaker_index), ivec(sc), ivec(min(ec, T)),
                                          return_masks=True)
        x_hat = ops.istft(X, self.stft_size, self.stft_shift, self.stft_fading)[0]
        if debug:
            masks = ops.unpack_fkt_to_ktf(post)[0].cpu().numpy().astype(np.float64)
            if self.bf_drop_context:
                masks[:, :sc, :] = 0
                if ec > 0:
                    masks[:, -ec:, :] = 0
            self.enhance_observation_locals = dict(
                acitivity_freq=acitivity_freq, masks=masks,
                target_mask=masks[target_speaker_index],
                distortion_mask=np.sum(np.delete(masks, target_speaker_index, axis=0), axis=0),
                X_hat=ops.unpack_ft_to_tf(X)[0].cpu().numpy(), x_hat=x_hat.cpu().numpy())
        return _from_device(x_hat, was_np, np.float64)

    def prepare_observation(self, obs, ex_array_activity, speaker_id, ex=None, pin=True, upload=True):
        """Host-side half of `enhance_observation` for one utterance, safe to run in a loader
        thread while the GPU works on the previous batch: float32 samples in page-locked memory,
        uploaded on a side stream (`upload`; the enhancement waits on the recorded event, so the
        host->device copy overlaps the kernels of the previous batch), frame-level activity (a4:
        boolean bookkeeping, database.py:409-472), target index and context frames.  core.py:514-547."""
        on_device = isinstance(obs, torch.Tensor) and obs.is_cuda
        ready = host = None
        if on_device:
            x = obs.to(torch.float32)
        else:
            x = obs if isinstance(obs, torch.Tensor) else torch.from_numpy(np.asarray(obs))
            if pin and torch.cuda.is_available() and not x.is_pinned():
                # page-locked staging buffers are recycled (page-locking 50 MB per utterance costs more
                # than the copy and serialises the ranks of a node in the kernel): a buffer is free again
                # once the upload that read it has completed
                src, x = x, self._pinned_buffer(tuple(x.shape))
                # NumPy copies a strided (D, N) cut row by row with memcpy (4x faster than torch's
                # strided copy_) and converts float64 -> float32 on the way
                np.copyto(x.numpy(), src.numpy(), casting='same_kind')
            else:
                x = x.to(torch.float32).contiguous()
            if upload and x.is_pinned():
                with self._pool_lock():
                    side = self.__dict__.get('_upload_stream')
                    if side is None or side.device.index != torch.cuda.current_device():
                        side = self.__dict__['_upload_stream'] = torch.cuda.Stream()
                with torch.cuda.stream(side):
                    host = x                                    # keep the pinned source alive until the copy is done
                    x = host.to(_device(), non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(side)
        assert x.ndim == 2, x.shape
        N = int(x.shape[-1])
        frames = ops.stft_frames(N, self.stft_size, self.stft_shift, self.stft_fading)
        af = activity_time_to_frequency(
            np.array([np.asarray(v) for v in ex_array_activity.values()]), stft_window_length=self.stft_size,
            stft_shift=self.stft_shift, stft_fading=self.stft_fading, stft_pad=True)
        sc = ec = 0
        if self.bf_drop_context and ex is not None:
            sc, ec = self._context_frames(ex)
        if host is not None:
            with self._pool_lock():
                self.__dict__.setdefault('_pinned_busy', []).append((ready, host))
        return dict(obs=x, ready=ready, host_source=host, N=N, frames=frames, activity_freq=af.astype(np.uint8), K=len(ex_array_activity),
                    target=tuple(ex_array_activity.keys()).index(speaker_id), start_ctx=sc, end_ctx=min(ec, frames),
                    numpy=not isinstance(obs, torch.Tensor))

    def _pool_lock(self):
        lock = self.__dict__.get('_pinned_lock')
        if lock is None:
            import threading
            lock = self.__dict__.setdefault('_pinned_lock', threading.Lock())
        return lock

    def _pinned_buffer(self, shape):
        """a float32 page-locked tensor of `shape` from the recycling pool (loader threads)"""
        n = int(np.prod(shape))
        flat = None
        with self._pool_lock():
            busy = self.__dict__.setdefault('_pinned_busy', [])     # (event, buffer) of uploads in flight
            free = self.__dict__.setdefault('_pinned_free', [])
            still = []
            for ev, buf in busy:
                if ev is None or ev.query():
                    free.append(buf._base if buf._base is not None else buf)
                else:
                    still.append((ev, buf))
            busy[:] = still
            best = None
            for i, b in enumerate(free):
                if b.numel() >= n and (best is None or b.numel() < free[best].numel()):
                    best = i
            if best is not None:
                flat = free.pop(best)
            while len(free) > 48:                                   # bound the pool
                free.pop(0)
        if flat is None:
            # capacities in steps of 16 MB: segments of similar length share buffers (page-locking is the
            # expensive part -- ~20 ms per 50 MB and serialised across the processes of a node)
            cap = -(-max(n, 1) // (1 << 22)) * (1 << 22)
            flat = torch.empty(cap, dtype=torch.float32, pin_memory=True)
        return flat[:n].view(shape)

    def enhance_prepared_batch(self, preps):
        """B prepared utterances of different lengths in ONE pass of the hot path: zero padded to
        the longest, frames beyond an utterance's own count masked through the ragged-batch
        interface, iSTFT output cut back.  Copies: one asynchronous host->device copy per utterance
        (pinned source), one device->host copy of the padded result."""
        B = len(preps)
        dev = _device()
        D, K = int(preps[0]['obs'].shape[0]), preps[0]['K']
        Ns = [p['N'] for p in preps]
        frames = [p['frames'] for p in preps]
        pad = (self.stft_size - self.stft_shift) if self.stft_fading else 0
        x = torch.zeros((B, D, max(Ns)), dtype=torch.float32, device=dev)
        cur = torch.cuda.current_stream()
        for b, p in enumerate(preps):
            assert p['obs'].shape[0] == D and p['K'] == K, (p['obs'].shape, D, p['K'], K)
            if p.get('ready') is not None:                      # uploaded by the loader thread on its side stream
                cur.wait_event(p['ready'])
                p['obs'].record_stream(cur)
            x[b, :, :Ns[b]].copy_(p['obs'], non_blocking=True)
        Y = ops.stft(x, self.stft_size, self.stft_shift, self.stft_fading)              # (B,F,D,Tmax)
        Tmax = Y.shape[3]
        act = np.zeros((B, K, Tmax), dtype=np.uint8)
        for b, p in enumerate(preps):
            n = min(p['activity_freq'].shape[-1], Tmax)
            act[b, :, :n] = p['activity_freq'][:, :n]
        ivec = lambda v: torch.tensor(v, dtype=torch.int32, device=dev)                 # noqa: E731
        info = ops.new_info(B, dev, stages=3)
        X = self.enhance_stft_batch(Y, torch.from_numpy(act).to(dev, non_blocking=True), ivec([p['target'] for p in preps]),
                                    ivec([p['start_ctx'] for p in preps]), ivec([p['end_ctx'] for p in preps]),
                                    frames=ivec(frames), info=info)
        x_hat = ops.istft(X, self.stft_size, self.stft_shift, self.stft_fading)        # (B, Nmax')
        n_out = [f * self.stft_shift + self.stft_size - self.stft_shift - 2 * pad for f in frames]
        if all(not p['numpy'] for p in preps):
            ops.check_info(info, self.STAGES)
            return [x_hat[b, :n_out[b]] for b in range(B)]
        flat = self.__dict__.get('_out_host')                    # grow-only page-locked result buffer; the returned
        if flat is None or flat.numel() < x_hat.numel():         # arrays are float64 copies, so it can be reused
            flat = self.__dict__['_out_host'] = torch.empty(x_hat.numel(), dtype=x_hat.dtype, pin_memory=True)
        host = flat[:x_hat.numel()].view(x_hat.shape)
        host.copy_(x_hat, non_blocking=True)
        info_h = info.cpu()                                   # synchronises: results and status words landed
        torch.cuda.current_stream().synchronize()
        ops.check_info(info_h, self.STAGES)
        hn = host.numpy()
        return [hn[b, :n_out[b]].astype(np.float64) if p['numpy'] else x_hat[b, :n_out[b]]
                for b, p in enumerate(preps)]

    def enhance_observation_batch(self, obs_list, ex_array_activities, speaker_ids, exs=None):
        """B utterances of different lengths in ONE pass of the hot path (what
        `enhance_observation` does per utterance, core.py:514-571).  Each result is bit-identical
        to the single-utterance call.  obs_list: B arrays (D, N_b) with the same D;
        ex_array_activities: B dicts {speaker: (N_b,) bool} with the same number of classes;
        exs: B example dicts (context frames) or None.  Returns a list of (N'_b,) arrays."""
        B = len(obs_list)
        assert B > 0 and len(ex_array_activities) == B and len(speaker_ids) == B
        preps = [self.prepare_observation(obs_list[b], ex_array_activities[b], speaker_ids[b],
                                          None if exs is None else exs[b]) for b in range(B)]
        return self.enhance_prepared_batch(preps)

    # ---- data plumbing (out of the hot path; uses the reference's I/O code) --
    def get_iterator(self, session_id):
        return self.db.get_iterator_for_session(
            session_id, audio_read=False, adjust_times=True, drop_unknown_target_speaker=True,
            context_samples=self.context_samples, equal_start_context=True)

    def enhance_session(self, session_ids, audio_dir, dataset_slice=False, audio_dir_exist_ok=False,
                        batch_size=8, skip_existing=False, strict=True, schedule='auto'):
        """core.py:333-394.  Work distribution: one process per GPU; rank r of W takes a static
        shard of the examples (the task farm of dlp_mpi.split_managed, see sharding.py) and runs
        it through the batching session driver (session.py: length-bucketed batches, audio
        prefetch, asynchronous wav writing).  `strict` (default, like the reference): the first
        failing example raises; `skip_existing`: resume an interrupted run."""
        from . import sharding
        from .session import SessionScheduler, run_distributed

        audio_dir = Path(audio_dir)
        it = self.get_iterator(session_ids)
        rank, world = sharding.init_process_group()      # binds the GPU of this rank (torchrun / mpiexec / srun)
        datasets = _session_to_dataset()
        if rank == 0:
            audio_dir.mkdir(exist_ok=audio_dir_exist_ok or skip_existing)
            for dataset in set(datasets.values()):
                (audio_dir / dataset).mkdir(exist_ok=audio_dir_exist_ok or skip_existing)
        sharding.barrier()
        if dataset_slice is not False:
            if dataset_slice is True:
                it = it[:2]
            elif isinstance(dataset_slice, int):
                it = it[:dataset_slice]
            elif isinstance(dataset_slice, slice):
                it = it[dataset_slice]
            else:
                raise ValueError(dataset_slice)
        examples = [it[i] for i in range(len(it))]

        def path_fn(ex):
            return audio_dir / datasets.get(ex['session_id'], 'unknown') / f'{ex["example_id"]}.wav'

        sched = SessionScheduler(self, self._load_example, path_fn, self._finish_example,
                                 batch_size=batch_size, skip_existing=skip_existing, strict=strict)
        return run_distributed(sched, examples, schedule)

    def _reference_array(self, ex):
        reference_array = self.reference_array
        if reference_array is None:
            try:
                reference_array = ex['reference_array']
            except KeyError:
                raise RuntimeError(
                    'Failed to get the "reference_array" from the example.\n'
                    'Probably you tried to enhance the "train" dataset.\n'
                    'Train has no "reference_array".\n'
                    'You can set a "reference_array" from the commandline with\n'
                    '\tpython -m ... with ... reference_array=U06\n'
                    'In case of multiarray, the reference array is used for the'
                    'projection of the human annotations.') from None
        return reference_array

    # -- example-dict layout (CHiME-5 json: per-array sample indices; core_chime6.py overrides) --
    def _context_frames(self, ex):
        return start_end_context_frames(ex, stft_size=self.stft_size, stft_shift=self.stft_shift,
                                        stft_fading=self.stft_fading)

    def _bounds(self, ex, array):
        """(start, stop) samples of the segment (with context) in `array`'s recording"""
        return ex['start']['observation'][array], ex['end']['observation'][array]

    def _orig(self, ex, reference_array):
        """(start_orig, num_samples_orig) of the utterance without context"""
        return (ex['start_orig']['observation'][reference_array],
                ex['num_samples_orig']['observation'][reference_array])

    def _session_activity(self, ex, reference_array):
        return self.activity[ex['session_id']][reference_array]

    def _load_example(self, ex):
        """core.py:396-498: slice the activity, load the audio of the selected arrays ->
        (obs (D, N) float64, ex_array_activity {speaker: (N,) bool}, speaker_id)."""
        from .audio_io import load_audio
        from .session import stack_arrays

        reference_array = self._reference_array(ex) if self._needs_reference_array() else None
        array_start, array_end = self._bounds(ex, reference_array)
        ex_array_activity = {
            k: arr[array_start:min(array_end, len(arr))]
            for k, arr in self._session_activity(ex, reference_array).items()}

        def load(array):
            start, stop = self._bounds(ex, array)
            x = load_audio(ex['audio_path']['observation'][array], start=start, stop=stop)
            return x[None] if x.ndim == 1 else x

        if self.multiarray is False:
            obs = load(self._reference_array(ex))
        else:
            obs = stack_arrays([load(a) for a in sorted(ex['audio_path']['observation'].keys())],
                               self.multiarray)                              # 'ACN->A*CN'
        return obs, ex_array_activity, ex['speaker_id']

    def _needs_reference_array(self):
        return True          # CHiME-5: activity and sample indices are per array

    def _finish_example(self, ex, x_hat):
        """cut the context again (core.py:500-505)"""
        if self.context_samples > 0:
            reference_array = self._reference_array(ex) if self._needs_reference_array() else None
            start_orig, num_samples_orig = self._orig(ex, reference_array)
            start, _ = self._bounds(ex, reference_array)
            start_context = start_orig - start
            x_hat = x_hat[..., start_context:start_context + num_samples_orig]
        return x_hat

    def enhance_example(self, ex, debug=False):
        """core.py:396-512: slice the activity, load the audio of the selected arrays,
        enhance, cut the context."""
        obs, ex_array_activity, speaker_id = self._load_example(ex)
        x_hat = self.enhance_observation(obs, ex_array_activity=ex_array_activity,
                                         speaker_id=speaker_id, ex=ex, debug=debug)
        x_hat = self._finish_example(ex, x_hat)
        if debug:
            self.enhance_example_locals = dict(obs=obs, ex_array_activity=ex_array_activity, x_hat=x_hat)
        return x_hat


def _session_to_dataset():
    """session -> dataset directory names (constants of pb_chime5/mapping.py:1-297, taken from the
    reference package when it is importable; otherwise the CHiME-5 table restated here)."""
    try:
        from pb_chime5 import mapping
        return dict(mapping.session_to_dataset)
    except Exception:  # noqa: BLE001
        table = {'train': ['S03', 'S04', 'S05', 'S06', 'S07', 'S08', 'S12', 'S13', 'S16', 'S17', 'S18', 'S19',
                           'S20', 'S22', 'S23', 'S24'],
                 'dev': ['S02', 'S09'], 'eval': ['S01', 'S21']}
        return {s: d for d, ss in table.items() for s in ss}


def get_enhancer(
    multiarray=False,
    reference_array=None,
    context_samples=240000,

    wpe=True,
    wpe_tabs=10,
    wpe_delay=2,
    wpe_iterations=3,
    wpe_psd_context=0,

    activity_type='annotation',
    activity_path=None,
    activity_garbage_class=True,

    stft_size=1024,
    stft_shift=256,
    stft_fading=True,

    bss_iterations=20,
    bss_iterations_post=1,

    bf_drop_context=True,

    bf='mvdrSouden_ban',
    postfilter=None,

    database_path=str(JSON_PATH / 'chime5.json'),
):
    """Factory with the reference's keyword names and defaults (core.py:574-637);
    sacred builds its config from this signature (scripts/run.py:19-27)."""
    assert wpe is True or wpe is False, wpe
    assert activity_path is None or activity_type == 'path', (activity_path, activity_type)
    return Enhancer(
        multiarray=multiarray,
        reference_array=reference_array,
        context_samples=context_samples,
        wpe_block=WPE(taps=wpe_tabs, delay=wpe_delay, iterations=wpe_iterations,
                      psd_context=wpe_psd_context) if wpe else None,
        activity=Activity(type=activity_type, garbage_class=activity_garbage_class,
                          path=activity_path, database_path=database_path),
        gss_block=GSS(iterations=bss_iterations, iterations_post=bss_iterations_post, verbose=False),
        bf_drop_context=bf_drop_context,
        bf_block=Beamformer(type=bf, postfilter=postfilter),
        stft_size=stft_size,
        stft_shift=stft_shift,
        stft_fading=stft_fading,
    )


def signature_defaults():
    """{kwarg: default} of get_enhancer, as scripts/run.py:23-25 reads it."""
    return {k: v.default for k, v in inspect.signature(get_enhancer).parameters.items()}
