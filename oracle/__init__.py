"""CPU oracle for the CLIBD hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy + a small C file) of the two reference
algorithms this repo accelerates:

* ``loss_oracle``  -- the contrastive loss of
  ``/root/reference/bioscanclip/model/loss_func.py`` (``ContrastiveLoss`` :25-69,
  ``ClipLoss`` :110-201, ``construct_label_metrix`` :19-22), forward and analytic
  backward.  Parity PINNED: checked against the reference's own PyTorch
  implementation through the golden vectors in ``tests/golden/`` that
  ``oracle/gen_golden.py`` produced by importing the reference in the build
  container.
* ``knn_oracle``   -- the cosine nearest-neighbour retrieval and top-k accuracy of
  ``/root/reference/bioscanclip/util/util.py`` (``make_prediction`` :521-553,
  ``top_k_micro_accuracy`` :379-395, ``top_k_macro_accuracy`` :555-599,
  ``inference_and_print_result`` :601-700).  The search arithmetic lives in the
  third-party ``faiss`` (``requirements.txt:22`` pins ``faiss-gpu==1.7.2``; the
  code uses the CPU ``IndexFlatIP``), which is absent from /root/reference and
  from this image, and the reference has no tests or golden vectors for it:
  PARITY UNPINNED for the search (exact inner product + lowest-index tie-break
  as BASELINE.json's north_star fixes it); the accuracy functions are pinned
  against the reference's own python functions (extracted by source from
  util.py, see gen_golden.py).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product
(``clibd_b200``) never imports it and fails loudly without its CUDA library.
"""
