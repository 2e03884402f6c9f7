"""oracle/ — CPU restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this package, and only as the checker / the timed CPU baseline.  Nothing under
`lipreading_b200/` imports it; the product path fails loudly when the CUDA library is missing.

Parity status:
  * sequence half (encoder, masked log-softmax, CTC wrapper, train/eval step): PINNED against the
    reference's own modules imported in the build container (`oracle/ref_harness.py`), golden
    vectors committed under `tests/golden/` by `tests/golden/make_golden.py`.  The RNN and CTC
    arithmetic itself lives in torch (un-vendored; reference pinned torch 0.4.1, oracle runs torch
    2.11 CPU fp32).  allennlp's masked_log_softmax / masked_softmax / sort_batch_by_length are
    un-vendored and un-pinned in the reference (install.ubuntu.sh:23); restated from the 0.7/0.8
    release that was current at the reference's date -> "parity unpinned" w.r.t. allennlp.
  * vision half (pad rect, crop box, similarity, warp, restore, gathers, translate): restated from
    src/utils/data/face.py and src/models/face/prnet.py; the gather indices are pinned by the
    reference's own fixture files (uv_kpt_ind.txt, face_ind.txt -> tests/golden/).  skimage 0.14.1
    (estimate_transform, warp), dlib 19.16 (HOG box) and the PRNet weights are absent from the
    reference tree and from this image -> "parity unpinned" for those; tolerances are stated in
    the tests.
  * position-map CNN (rows a5 / f4): oracle/tapgemm.py states what one lr_tapgemm launch computes; tests replay the
    compiled launch plan through it and hold it to prnet.ResFcn256 (the restatement of src/models/face/prnet.py:211-280,
    itself held to the reference's checkpoint index).  No weights exist -> values "parity unpinned".
  * conv3d front-end and mouth crop are north-star extensions with no reference code; their oracle
    is torch.nn.functional.conv3d fp32 / the numpy spec in oracle/vision.py.
"""
