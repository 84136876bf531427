"""ORACLE — CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
anything under oracle/.  The product package (manual_whisper_b200/) never does and fails loudly when
its CUDA library is missing.

Parity status: the reference (/root/reference) holds no tests, golden vectors or fixtures for this path
and its arithmetic lives in un-vendored third-party packages (whisperx==3.7.6 ->
faster-whisper>=1.1.1 -> ctranslate2>=4.5) that are absent here, so the restatement follows their
published algorithm and is pinned against the independent Hugging Face
implementation that IS importable in this container (log-mel: bit-equal; encoder/decoder logits: 1e-4;
timestamp rules: equal masks).  Beam-search ordering has no second opinion: "parity unpinned" there.
"""
