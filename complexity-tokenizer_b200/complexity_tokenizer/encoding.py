"""`Encoding` / `BatchEncoding`: the Python face of the rich outputs computed on the GPU (csrc/encoding.cu).

Mirrors the reference's PyO3 classes (/root/reference/src/bindings/encoding.rs:7-167 `Encoding`, :170-296
`BatchEncoding`): same getters, methods and return types.  Every array of a batch is computed on the device and
arrives here packed; an `Encoding` is a window into those arrays.  What remains on the host is marshalling: turning
NumPy windows into Python lists, and the in-place `pad` / `truncate` of ONE object (encoding.rs:86-232), which only
appends constants or cuts the object's own arrays.
"""
import numpy as np


class Encoding:
    """bindings/encoding.rs:7-167.  Arrays are NumPy; getters return the Python types PyO3 returns."""

    def __init__(self, ids, type_ids, attention_mask, special_tokens_mask, tokens, offsets, word_ids, sequence_ids,
                 overflow_fn=None):
        self._ids = np.asarray(ids, dtype=np.uint32)
        self._type_ids = np.asarray(type_ids, dtype=np.uint32)
        self._attn = np.asarray(attention_mask, dtype=np.uint32)
        self._spec = np.asarray(special_tokens_mask, dtype=np.uint32)
        self._tokens = tokens                      # list[str], or a function that builds it on first use
        self._offsets = np.asarray(offsets, dtype=np.uint64).reshape(-1, 2)
        self._word_ids = np.asarray(word_ids, dtype=np.int64)
        self._seq = np.asarray(sequence_ids, dtype=np.int64)      # -1 = None
        self._overflow = []
        self._overflow_fn = overflow_fn            # builds the overflowing windows of a truncated row on first use

    @staticmethod
    def from_ids(ids, tokens):
        """encoding.rs:44-58"""
        n = len(ids)
        return Encoding(ids, np.zeros(n), np.ones(n), np.zeros(n), list(tokens), np.zeros((0, 2)), np.zeros(0), np.zeros(n))

    # ---- getters
    @property
    def ids(self):
        return self._ids.tolist()

    @property
    def tokens(self):
        if callable(self._tokens):
            self._tokens = self._tokens()
        return list(self._tokens)

    @property
    def attention_mask(self):
        return self._attn.tolist()

    @property
    def type_ids(self):
        return self._type_ids.tolist()

    @property
    def special_tokens_mask(self):
        return self._spec.tolist()

    @property
    def offsets(self):
        return [tuple(p) for p in self._offsets.tolist()]

    @property
    def word_ids(self):
        return self._word_ids.tolist()

    @property
    def sequence_ids(self):
        return [None if v < 0 else v for v in self._seq.tolist()]

    def __len__(self):
        return int(self._ids.size)

    def _ensure_overflow(self):
        if self._overflow_fn is not None:
            fn, self._overflow_fn = self._overflow_fn, None
            self._overflow = list(fn()) + self._overflow

    @property
    def n_overflowing(self):
        self._ensure_overflow()
        return len(self._overflow)

    @property
    def overflowing(self):
        self._ensure_overflow()
        return list(self._overflow)

    def ids_as_numpy(self):
        return self._ids.copy()

    def attention_mask_as_numpy(self):
        return self._attn.copy()

    def type_ids_as_numpy(self):
        return self._type_ids.copy()

    def special_tokens_mask_as_numpy(self):
        return self._spec.copy()

    # ---- in-place edits of one object (encoding.rs:86-232)
    def pad(self, target_length, pad_id, pad_token, pad_left):
        n = len(self)
        if n >= target_length:
            return
        k = target_length - n
        toks = self.tokens

        def ext(a, fill):
            f = np.full(k, fill, dtype=a.dtype)
            return np.concatenate([f, a]) if pad_left else np.concatenate([a, f])

        self._ids = ext(self._ids, pad_id)
        self._type_ids = ext(self._type_ids, 0)
        self._attn = ext(self._attn, 0)
        self._spec = ext(self._spec, 1)
        self._seq = ext(self._seq, -1)
        self._tokens = [pad_token] * k + toks if pad_left else toks + [pad_token] * k

    def _window(self, a, b):
        toks = self.tokens
        return Encoding(self._ids[a:b], self._type_ids[a:b], self._attn[a:b], self._spec[a:b], toks[a:b],
                        self._offsets[a:b] if len(self._offsets) > a else np.zeros((0, 2)),
                        self._word_ids[a:b] if len(self._word_ids) > a else np.zeros(0),
                        self._seq[a:b] if len(self._seq) > a else np.zeros(0))

    def _cut(self, n):
        toks = self.tokens
        self._ids, self._type_ids, self._attn, self._spec = self._ids[:n], self._type_ids[:n], self._attn[:n], self._spec[:n]
        self._tokens, self._offsets, self._word_ids, self._seq = toks[:n], self._offsets[:n], self._word_ids[:n], self._seq[:n]

    def truncate(self, max_length):
        if len(self) <= max_length:
            return
        self._ensure_overflow()
        self._overflow.append(self._window(max_length, len(self)))
        self._cut(max_length)

    def truncate_with_stride(self, max_length, stride):
        if len(self) <= max_length:
            return
        if stride >= max_length:
            raise ValueError('stride >= max_length: the reference loops forever (encoding.rs:190-193)')
        self._ensure_overflow()
        pos, n = max_length, len(self)
        while pos < n:
            start = max(pos - stride, 0)
            end = min(start + max_length, n)
            self._overflow.append(self._window(start, end))
            pos = end
        self._cut(max_length)

    # ---- look-ups (encoding.rs:270-390)
    def char_to_token(self, char_pos):
        for i, (s, e) in enumerate(self._offsets.tolist()):
            if s <= char_pos < e:
                return i
        return None

    def char_to_token_with_sequence(self, char_pos, sequence_id):
        seq = self._seq.tolist()
        for i, (s, e) in enumerate(self._offsets.tolist()):
            if i < len(seq) and seq[i] == sequence_id and s <= char_pos < e:
                return i
        return None

    def token_to_chars(self, token_idx):
        return tuple(self._offsets[token_idx].tolist()) if 0 <= token_idx < len(self._offsets) else None

    def token_to_word(self, token_idx):
        return int(self._word_ids[token_idx]) if 0 <= token_idx < len(self._word_ids) else None

    def token_to_sequence(self, token_idx):
        if 0 <= token_idx < len(self._seq) and self._seq[token_idx] >= 0:
            return int(self._seq[token_idx])
        return None

    def word_to_tokens(self, word_idx):
        return self.word_to_tokens_with_sequence(word_idx, 0)

    def word_to_tokens_with_sequence(self, word_idx, sequence_id=0):
        n = min(len(self._word_ids), len(self._seq))
        hit = np.nonzero((self._word_ids[:n] == word_idx) & (self._seq[:n] == sequence_id))[0]
        return (int(hit[0]), int(hit[-1]) + 1) if hit.size else None

    def word_to_chars(self, word_idx):
        return self.word_to_chars_with_sequence(word_idx, 0)

    def word_to_chars_with_sequence(self, word_idx, sequence_id=0):
        r = self.word_to_tokens_with_sequence(word_idx, sequence_id)
        if r is None:
            return None
        o = self._offsets[r[0]:r[1]]
        return (int(o[:, 0].min()), int(o[:, 1].max())) if len(o) else None

    def word_token_indices(self, word_idx):
        return np.nonzero(self._word_ids == word_idx)[0].tolist()

    @property
    def n_words(self):
        return int(self._word_ids.max()) + 1 if self._word_ids.size else 0


class PackedEncodings:
    """One batch as it comes back from ctk_encode_batch_to_encoding: packed NumPy arrays (copies; the C result is freed)."""

    def __init__(self, tokenizer, **kw):
        self.tokenizer = tokenizer
        self.__dict__.update(kw)

    @property
    def n_rows(self):
        return len(self.row_off) - 1

    def dense(self, name):
        """[n_rows, L] matrix of `input_ids` / `attention_mask` / `token_type_ids` / `special_tokens_mask` when every row
        has the same length (padding), else ValueError."""
        a = getattr(self, name)
        n = self.n_rows
        lens = np.diff(self.row_off)
        if n and not (lens == lens[0]).all():
            raise ValueError('rows have different lengths: ask for padding')
        return a.reshape(n, int(lens[0]) if n else 0)

    def encoding(self, r):
        g = 2 if self.pair else 1
        t0, t2 = int(self.tok_off[r * g]), int(self.tok_off[(r + 1) * g])
        n = t2 - t0
        len_a = (int(self.tok_off[r * g + 1]) - t0) if g == 2 else n
        a, b = int(self.row_off[r]), int(self.row_off[r + 1])
        full = int(self.row_full[r])
        cut = min(full, self.max_length) if self.truncation else full
        truncated = self.truncation and full > self.max_length
        npad = (b - a) - cut
        tok = self.tokenizer
        raw = self.raw_ids[t0:t2]
        keep = self.max_length if truncated else n
        seq = np.concatenate([np.zeros(len_a, dtype=np.int64), np.ones(n - len_a, dtype=np.int64)])[:keep]
        pad_left, pad_token = self.pad_left, self.pad_token

        def tokens():
            if self.add_special_tokens:
                toks = [tok.id_to_token(int(i)) or '' for i in raw.tolist()]
            else:                                   # bindings/tokenizer.rs:92: filter_map drops ids without a token
                toks = [t for t in (tok.id_to_token(int(i)) for i in raw.tolist()) if t is not None]
            toks = toks[:keep]
            return [pad_token] * npad + toks if pad_left else toks + [pad_token] * npad

        if npad:
            f = np.full(npad, -1, dtype=np.int64)
            seq = np.concatenate([f, seq]) if pad_left else np.concatenate([seq, f])
        if self.add_special_tokens:                 # from_ids leaves offsets and word ids empty (encoding.rs:52-53)
            all_offs, all_wids = self.offsets_and_word_ids()
            offs, wids = all_offs[t0:t2][:keep], all_wids[t0:t2][:keep].astype(np.int64)
        else:
            offs, wids = np.zeros((0, 2)), np.zeros(0)
        fn = (lambda: self.overflow_of(r)) if truncated else None
        return Encoding(self.input_ids[a:b], self.token_type_ids[a:b], self.attention_mask[a:b], self.special_tokens_mask[a:b],
                        tokens, offs, wids, seq, fn)

    def offsets_and_word_ids(self):
        """Per-token (start, end) byte offsets and word ids of the whole batch; computed on the device on first use (a batch
        that is only asked for input_ids / masks never pays for them)."""
        if self.offsets is None:
            q = self.tokenizer._encode_rows(self.text_buf, self.text_off, pair=self.pair, add_special_tokens=True, want_offsets=True)
            self.offsets, self.word_ids = q.offsets, q.word_ids
        return self.offsets, self.word_ids

    def reencode_row(self, r):
        g = 2 if self.pair else 1
        o = self.text_off[r * g:(r + 1) * g + 1]
        buf = self.text_buf[int(o[0]):int(o[-1])]
        q = self.tokenizer._encode_rows(buf, o - o[0], pair=self.pair, add_special_tokens=self.add_special_tokens,
                                        want_offsets=self.add_special_tokens)
        return q.encoding(0)

    def overflow_of(self, r):
        """The overflowing windows of a truncated row (encoding.rs:139-165, :190-218): windows of the same row encoded
        without truncation."""
        full = self.reencode_row(r)
        n, m = len(full), self.max_length
        if self.overflow_mode == 'single':
            return [full._window(m, n)]
        out, pos = [], m
        while pos < n:
            start = max(pos - self.stride, 0)
            end = min(start + m, n)
            out.append(full._window(start, end))
            pos = end
        return out


class BatchEncoding:
    """bindings/encoding.rs:170-296 (the result of `tokenizer(texts, ...)`)."""

    def __init__(self, packed, return_attention_mask=True, return_token_type_ids=True, return_offsets_mapping=False,
                 return_special_tokens_mask=False):
        self._p = packed
        self._attn, self._type, self._offs, self._spec = (return_attention_mask, return_token_type_ids,
                                                          return_offsets_mapping, return_special_tokens_mask)

    def _rows(self, a):
        o = self._p.row_off.tolist()
        lst = a.tolist()
        return [lst[o[i]:o[i + 1]] for i in range(len(o) - 1)]

    @property
    def input_ids(self):
        return self._rows(self._p.input_ids)

    @property
    def attention_mask(self):
        return self._rows(self._p.attention_mask) if self._attn else []

    @property
    def token_type_ids(self):
        return self._rows(self._p.token_type_ids) if self._type else []

    @property
    def special_tokens_mask(self):
        return self._rows(self._p.special_tokens_mask) if self._spec else []

    @property
    def offset_mapping(self):
        if not self._offs:
            return []
        return [self._p.encoding(r).offsets for r in range(len(self))]

    def encodings(self):
        return [self._p.encoding(r) for r in range(len(self))]

    def __len__(self):
        return self._p.n_rows

    def __getitem__(self, idx):
        if not 0 <= idx < len(self):
            raise IndexError('Index out of range')
        return self._p.encoding(idx)

    def keys(self):
        k = ['input_ids']
        if self._attn:
            k.append('attention_mask')
        if self._type:
            k.append('token_type_ids')
        if self._spec:
            k.append('special_tokens_mask')
        if self._offs:
            k.append('offset_mapping')
        return k

    def input_ids_as_numpy(self):
        o = self._p.row_off.tolist()
        return [self._p.input_ids[o[i]:o[i + 1]].copy() for i in range(len(o) - 1)]

    def attention_mask_as_numpy(self):
        o = self._p.row_off.tolist()
        return [self._p.attention_mask[o[i]:o[i + 1]].astype(np.uint32) for i in range(len(o) - 1)]

    def to_dict(self):
        return {k: getattr(self, k) for k in self.keys()}

    # not in the reference: the batch as dense [rows, L] matrices without building Python lists (needs padding)
    def as_numpy(self, name='input_ids'):
        return self._p.dense(name)
