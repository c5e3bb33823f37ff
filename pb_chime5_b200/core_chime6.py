"""CHiME-6 front door: the same blocks and kernels behind the example layout of
``pb_chime5/core_chime6.py`` (row f4 of SURVEY.md section 8).

The reference's ``core_chime6.py`` is a copy of ``core.py`` that differs only in plumbing
(``diff core.py core_chime6.py``): the CHiME-6 json stores ONE synchronised sample index per
example (``ex['start']``, ``ex['end']``, ``ex['start_orig']``, ``ex['num_samples_orig']`` are
plain ints, core_chime6.py:217-218,403-404,484-502), the activity is per session instead of per
array (``get_activity_chime6``, :113-118; ``self.activity[session_id]``, :408), the iterator is
built with ``adjust_times=False, equal_start_context=False`` (:326-329), the reference array is
only needed to pick the recording when ``multiarray`` is off (:468-485), and the database is
``chime6.json`` (:97,599).  ``WPE``, ``GSS`` and ``Beamformer`` are the classes of
``pb_chime5_b200.core`` (the reference re-uses them the same way, core_chime6_rttm.py:26).
"""
from __future__ import annotations

import inspect
from dataclasses import dataclass

from . import core as _core
from .core import GSS, WPE, Beamformer, JSON_PATH, samples_to_stft_frames  # noqa: F401  (re-exported API)


@dataclass
class Activity(_core.Activity):
    """core_chime6.py:91-140"""
    database_path: str = str(JSON_PATH / 'chime6.json')

    def __getitem__(self, session_id):
        if self.type == 'annotation':
            return _annotation_activity_chime6(session_id, self.db, self.garbage_class)
        return super().__getitem__(session_id)


_ACTIVITY_CACHE = {}


def _annotation_activity_chime6(session_id, db, garbage_class):
    key = (session_id, id(db), garbage_class)
    if key not in _ACTIVITY_CACHE:
        from pb_chime5.activity import get_activity_chime6   # reference metadata code
        _ACTIVITY_CACHE.clear()
        _ACTIVITY_CACHE[key] = get_activity_chime6(
            iterator=db.get_datasets(session_id), garbage_class=garbage_class, dtype=bool,
            non_sil_alignment_fn=None, debug=False, use_ArrayIntervall=True)[session_id]
    return _ACTIVITY_CACHE[key]


def start_end_context_frames(ex, stft_size, stft_shift, stft_fading):
    """core_chime6.py:216-237 (flat sample indices)"""
    start_context_samples = ex['start_orig'] - ex['start']
    end_context_samples = ex['end'] - ex['end_orig']
    assert start_context_samples >= 0, (start_context_samples, ex)
    assert end_context_samples >= 0, (end_context_samples, ex)
    return (samples_to_stft_frames(start_context_samples, stft_size, stft_shift, fading=stft_fading),
            samples_to_stft_frames(end_context_samples, stft_size, stft_shift, fading=stft_fading))


@dataclass
class Enhancer(_core.Enhancer):
    """core_chime6.py:280-570: same pipeline, CHiME-6 example layout."""

    def get_iterator(self, session_id):
        return self.db.get_iterator_for_session(
            session_id, audio_read=False, adjust_times=False, drop_unknown_target_speaker=True,
            context_samples=self.context_samples, equal_start_context=False)

    def _context_frames(self, ex):
        return start_end_context_frames(ex, stft_size=self.stft_size, stft_shift=self.stft_shift,
                                        stft_fading=self.stft_fading)

    def _needs_reference_array(self):
        return False

    def _bounds(self, ex, array):
        return ex['start'], ex['end']

    def _orig(self, ex, reference_array):
        return ex['start_orig'], ex['num_samples_orig']

    def _session_activity(self, ex, reference_array):
        return self.activity[ex['session_id']]


def get_enhancer(
    multiarray=False,
    context_samples=240000,
    reference_array=None,

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

    database_path=str(JSON_PATH / 'chime6.json'),
):
    """core_chime6.py:573-635 (keyword names, order and defaults are API: sacred reads them)."""
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
    return {k: v.default for k, v in inspect.signature(get_enhancer).parameters.items()}
