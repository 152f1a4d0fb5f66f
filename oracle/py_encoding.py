"""ORACLE (test infrastructure, not product code) -- rich `Encoding` outputs, pure Python.

CPU restatement of the reference's `Encoding` path, line by line.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import this; the product never does.

PARITY PINNING: same status as oracle/py_oracle.py (the Rust reference cannot be built here).  Pinned on the
known answers of the reference's own unit tests: src/encoding.rs:464-576 (from_ids, pad, truncate, char/word
look-ups) and src/postprocessors.rs:295-356 (Bert/Roberta processing, pad); see tests/test_encoding_oracle.py.

Reference lines restated here (paths relative to /root/reference):
  Encoding            src/encoding.rs:7-26, :44-58 (from_ids), :77-83 (mark_special_tokens), :86-129 (pad),
                      :132-180 (truncate), :183-232 (truncate_with_stride), :250-267 (merge)
  post-processors     src/postprocessors.rs:34-55 (process), :88-148 (template walk), :151-188 (Bert, Roberta)
                      src/huggingface/parsing.rs:193-250 (parse_post_processor), :253-270 (template_from_array)
  encode_to_encoding  src/huggingface/mod.rs:340-395 (impl), :397-444 (single), :447-478 (word offsets),
                      :481-545 (batch, padding variants)
  __call__            src/bindings/tokenizer.rs:33-201

Everything is byte arithmetic on UTF-8: "offsets" are BYTE positions in the original text (mod.rs:461-475 index
`original` with `str::find` results), and a token's length is the byte length of its vocabulary string.
"""
import unicodedata as ud

from py_oracle import ReferencePanic


class Enc:
    """src/encoding.rs:7-26"""

    def __init__(self):
        self.ids = []
        self.type_ids = []
        self.tokens = []
        self.attention_mask = []
        self.special_tokens_mask = []
        self.offsets = []
        self.word_ids = []
        self.sequence_ids = []
        self.overflowing = []

    @classmethod
    def from_ids(cls, ids, tokens):
        """encoding.rs:44-58"""
        e = cls()
        n = len(ids)
        e.ids = list(ids)
        e.type_ids = [0] * n
        e.tokens = list(tokens)
        e.attention_mask = [1] * n
        e.special_tokens_mask = [0] * n
        e.sequence_ids = [0] * n
        return e

    def __len__(self):
        return len(self.ids)

    def mark_special_tokens(self, special_ids):
        """encoding.rs:77-83"""
        s = set(special_ids)
        for i, t in enumerate(self.ids):
            if t in s:
                self.special_tokens_mask[i] = 1

    def pad(self, target, pad_id, pad_token, pad_left):
        """encoding.rs:86-129: offsets and word_ids are NOT padded"""
        if len(self) >= target:
            return
        k = target - len(self)
        if pad_left:
            self.ids = [pad_id] * k + self.ids
            self.type_ids = [0] * k + self.type_ids
            self.tokens = [pad_token] * k + self.tokens
            self.attention_mask = [0] * k + self.attention_mask
            self.special_tokens_mask = [1] * k + self.special_tokens_mask
            self.sequence_ids = [None] * k + self.sequence_ids
        else:
            self.ids += [pad_id] * k
            self.type_ids += [0] * k
            self.tokens += [pad_token] * k
            self.attention_mask += [0] * k
            self.special_tokens_mask += [1] * k
            self.sequence_ids += [None] * k

    def _window(self, a, b):
        o = Enc()
        o.ids = self.ids[a:b]
        o.type_ids = self.type_ids[a:b]
        o.tokens = self.tokens[a:b]
        o.attention_mask = self.attention_mask[a:b]
        o.special_tokens_mask = self.special_tokens_mask[a:b]
        o.offsets = self.offsets[a:b] if len(self.offsets) > a else []
        o.word_ids = self.word_ids[a:b] if len(self.word_ids) > a else []
        o.sequence_ids = self.sequence_ids[a:b] if len(self.sequence_ids) > a else []
        return o

    def _cut(self, n):
        for f in ('ids', 'type_ids', 'tokens', 'attention_mask', 'special_tokens_mask', 'offsets', 'word_ids',
                  'sequence_ids'):
            setattr(self, f, getattr(self, f)[:n])

    def truncate(self, max_length):
        """encoding.rs:132-180"""
        if len(self) <= max_length:
            return
        self.overflowing.append(self._window(max_length, len(self)))
        self._cut(max_length)

    def truncate_with_stride(self, max_length, stride):
        """encoding.rs:183-232"""
        if len(self) <= max_length:
            return
        pos = max_length
        guard = 0
        while pos < len(self.ids):
            start = max(pos - stride, 0)
            end = min(start + max_length, len(self.ids))
            self.overflowing.append(self._window(start, end))
            if end <= pos:
                guard += 1
                if guard > 4:
                    raise ReferencePanic('truncate_with_stride does not advance (stride >= max_length): the reference loops forever')
            pos = end
        self._cut(max_length)

    def merge(self, other, type_id):
        """encoding.rs:250-267"""
        n = len(other.ids)
        self.ids += other.ids
        self.tokens += other.tokens
        self.attention_mask += other.attention_mask
        self.special_tokens_mask += other.special_tokens_mask
        self.offsets += other.offsets
        self.word_ids += other.word_ids
        self.type_ids += [type_id] * n
        self.sequence_ids += [type_id] * n


# ------------------------------------------------------------------------------------------------
# post-processors

def template_from_array(arr):
    """parsing.rs:253-270"""
    parts = []
    for item in arr:
        if isinstance(item, dict):
            sp = item.get('SpecialToken')
            if sp is not None:
                v = sp.get('id') if isinstance(sp, dict) else None
                if isinstance(v, str):
                    parts.append(v)
                continue
            seq = item.get('Sequence')
            if seq is not None:
                v = seq.get('id') if isinstance(seq, dict) else None
                if isinstance(v, str):
                    parts.append('$' + v)
    return ' '.join(parts)


def parse_post_processor(v, special_tokens):
    """parsing.rs:193-250 -> ('template', single, pair, specials) | ('roberta', bos, eos) | ('bert', cls, sep) | None"""
    if isinstance(v, dict) and 'type' in v:
        t = v['type'] if isinstance(v['type'], str) else ''
        if t == 'TemplateProcessing':
            s = v.get('single')
            single = template_from_array(s) if isinstance(s, list) else '<s> $A </s>'
            p = v.get('pair')
            pair = template_from_array(p) if isinstance(p, list) else None
            return ('template', single, pair, dict(special_tokens))
        if t == 'RobertaProcessing':
            return ('roberta', special_tokens.get('<s>', 0), special_tokens.get('</s>', 2))
        if t == 'BertProcessing':
            return ('bert', special_tokens.get('[CLS]', 101), special_tokens.get('[SEP]', 102))
    return None


def template_process(ids, pair_ids, single, pair, specials):
    """postprocessors.rs:88-148"""
    template = (pair if pair is not None else single) if pair_ids is not None else single
    chars = list(template)
    out = []
    i = 0
    while i < len(chars):
        if chars[i] == '$' and i + 1 < len(chars):
            if chars[i + 1] == 'A':
                out += ids
                i += 2
            elif chars[i + 1] == 'B':
                if pair_ids is not None:
                    out += pair_ids
                i += 2
            else:
                i += 1
        elif chars[i] in '<[':
            endc = '>' if chars[i] == '<' else ']'
            start = i
            while i < len(chars) and chars[i] != endc:
                i += 1
            if i < len(chars):
                i += 1
            tok = ''.join(chars[start:i]).strip()      # Rust str::trim = White_Space
            if tok in specials:
                out.append(specials[tok])
        else:
            i += 1
    return out


def post_process(pp, ids, pair_ids=None):
    """postprocessors.rs:34-55 (the Sequence arm cannot come out of parse_post_processor)"""
    if pp[0] == 'template':
        return template_process(ids, pair_ids, pp[1], pp[2], pp[3])
    if pp[0] == 'bert':            # :151-166
        out = [pp[1]] + ids + [pp[2]]
        if pair_ids is not None:
            out += pair_ids + [pp[2]]
        return out
    out = [pp[1]] + ids + [pp[2]]  # roberta :169-188
    if pair_ids is not None:
        out += [pp[2]] + pair_ids + [pp[2]]
    return out


# ------------------------------------------------------------------------------------------------
# the tokenizer-level methods

class RichOracle:
    """Methods of HuggingFaceTokenizer / PyTokenizer that return Encoding objects, over an OracleTokenizer."""

    def __init__(self, tok, tokenizer_json):
        self.tok = tok
        self.post_processor = parse_post_processor(tokenizer_json.get('post_processor'), tok.special_tokens)
        self.model_max_length = 512            # mod.rs:243-245
        self.padding_side = 'right'            # mod.rs:325

    def _token_str(self, i):
        s = self.tok.id_to_token.get(i)
        return s if s is not None else ''

    def pre_tokenize_with_offsets(self, normalized, original):
        """mod.rs:447-478; positions are byte indices into `original`"""
        ob = original.encode('utf-8')
        out = []
        ss = 0
        for word in self.tok.pre_tokenize(normalized):
            trimmed = word.lstrip('Ġ▁')
            to_find = (trimmed if trimmed else word).encode('utf-8')
            if ss < len(ob) and (ob[ss] & 0xC0) == 0x80:
                raise ReferencePanic('byte index %d is not a char boundary' % ss)    # original[search_start..]
            pos = ob.find(to_find, ss)
            if pos >= 0:
                start, end = pos, pos + len(to_find)
            else:
                start = ss
                end = min(start + len(word.encode('utf-8')), len(ob))
            out.append((word, start, end))
            ss = end
        return out

    def encode_single_to_encoding(self, text, type_id):
        """mod.rs:397-444 (no added-token scan on this path)"""
        norm = ud.normalize('NFC', text) if self.tok.normalizer == 'nfc' else text
        e = Enc()
        for wi, (word, ws, we) in enumerate(self.pre_tokenize_with_offsets(norm, text)):
            off = ws
            for i in self.tok.bpe_encode(word):
                s = self._token_str(i)
                end = min(off + len(s.encode('utf-8')), we)
                e.ids.append(i)
                e.offsets.append((off, end))
                off = end
                e.tokens.append(s)
                e.word_ids.append(wi)
        n = len(e.ids)
        e.type_ids = [type_id] * n
        e.attention_mask = [1] * n
        e.special_tokens_mask = [0] * n
        e.sequence_ids = [type_id] * n
        return e

    def encode_to_encoding(self, text, text_pair=None, max_length=None, stride=None):
        """mod.rs:357-395"""
        e = self.encode_single_to_encoding(text, 0)
        if text_pair is not None:
            e.merge(self.encode_single_to_encoding(text_pair, 1), 1)
        processed = post_process(self.post_processor, list(e.ids)) if self.post_processor else list(e.ids)
        if len(processed) < len(e.ids):
            raise ReferencePanic('attempt to subtract with overflow (mod.rs:377)')
        added = len(processed) - len(e.ids)
        e.ids = processed
        e.attention_mask += [1] * added
        e.special_tokens_mask += [1] * added
        e.type_ids += [0] * added
        e.mark_special_tokens(self.tok.special_tokens.values())
        if max_length is not None and len(e) > max_length:
            e.truncate_with_stride(max_length, stride or 0)
        return e

    def _from_ids(self, text):
        """bindings/tokenizer.rs:88-96: the filter_map drops tokens of ids that are not in the vocabulary"""
        ids = self.tok.encode(text)
        return Enc.from_ids(ids, [self.tok.id_to_token[i] for i in ids if i in self.tok.id_to_token])

    def _pad_id_token(self, via_vocab_default):
        sp = self.tok.special_tokens
        pad_id = sp.get('[PAD]', sp.get('<pad>', 0))
        tok = self.tok.id_to_token.get(pad_id)
        return pad_id, (tok if tok is not None else '<pad>')

    def encode_batch_with_padding(self, texts, pad_to_max=None, pad_left=False, pairs=False):
        """mod.rs:490-545"""
        encs = [self.encode_to_encoding(*t) if pairs else self.encode_to_encoding(t) for t in texts]
        max_len = pad_to_max if pad_to_max is not None else max([len(e) for e in encs], default=0)
        pad_id, pad_token = self._pad_id_token(True)
        for e in encs:
            e.pad(max_len, pad_id, pad_token, pad_left)
        return encs

    def call(self, text, text_pair=None, add_special_tokens=True, padding=None, truncation=False, max_length=None,
             stride=0):
        """bindings/tokenizer.rs:46-201 -> list of Enc (the BatchEncoding's encodings)"""
        single = isinstance(text, str)
        texts = [text] if single else list(text)
        if single:
            pairs = [text_pair] if isinstance(text_pair, str) else None
        else:
            pairs = list(text_pair) if isinstance(text_pair, (list, tuple)) else None
        encs = []
        for k, t in enumerate(texts if pairs is None else list(zip(texts, pairs))):
            if pairs is not None:
                a, b = t
                if add_special_tokens:
                    encs.append(self.encode_to_encoding(a, b))
                else:
                    e = self._from_ids(a)
                    e.merge(self._from_ids(b), 1)
                    encs.append(e)
            else:
                encs.append(self.encode_to_encoding(t) if add_special_tokens else self._from_ids(t))
        max_len = max_length if max_length is not None else self.model_max_length
        if truncation:
            for e in encs:
                if len(e) > max_len:
                    if stride > 0:
                        e.truncate_with_stride(max_len, stride)
                    else:
                        e.truncate(max_len)
        if padding is not None:
            if padding == 'max_length':
                target = max_len
            else:
                target = max([len(e) for e in encs], default=0)
            pad_id, pad_token = self._pad_id_token(False)
            pad_left = padding == 'left' or self.padding_side == 'left'
            for e in encs:
                e.pad(target, pad_id, pad_token, pad_left)
        return encs
