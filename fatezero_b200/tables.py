"""Host-side edit tables, computed once per edit (tiny integer / fp32 work, stays on the CPU like in the reference):
word -> token indices, per-step cross-replace alpha, refinement mapper (global alignment of the two token
sequences), replacement matrix, equalizer.  Semantics of prompt_attention/ptp_utils.py:144-199,
prompt_attention/seq_aligner.py:61-195 and prompt_attention/attention_util.py:307-316."""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple, Union

import numpy as np
import torch

MAX_WORDS = 77


def get_word_inds(text: str, word_place: Union[int, str, Sequence[int]], tokenizer) -> np.ndarray:
    """Token positions (BOS = position 0) covered by the selected word(s) of `text`."""
    words = text.split(" ")
    if isinstance(word_place, str):
        wanted = {i for i, w in enumerate(words) if w == word_place}
    elif isinstance(word_place, int):
        wanted = {word_place}
    else:
        wanted = set(int(i) for i in word_place)
    hits: List[int] = []
    if wanted:
        pieces = [tokenizer.decode([tid]).strip("#") for tid in tokenizer.encode(text)][1:-1]
        word_idx, consumed = 0, 0
        for pos, piece in enumerate(pieces, start=1):
            consumed += len(piece)
            if word_idx in wanted:
                hits.append(pos)
            if consumed >= len(words[word_idx]):
                word_idx, consumed = word_idx + 1, 0
    return np.array(hits)


def get_time_words_attention_alpha(prompts: Sequence[str], num_steps: int, cross_replace_steps, tokenizer,
                                   max_num_words: int = MAX_WORDS) -> torch.Tensor:
    """[num_steps+1, len(prompts)-1, 1, 1, 77]: 1 while the cross-attention of a token is taken from the source."""
    spec: Dict[str, Union[float, Tuple[float, float]]]
    if hasattr(cross_replace_steps, "items"):
        spec = dict(cross_replace_steps.items())
    else:
        spec = {"default_": cross_replace_steps}
    spec.setdefault("default_", (0.0, 1.0))
    n_t = num_steps + 1
    table = torch.zeros(n_t, len(prompts) - 1, max_num_words)

    def paint(bounds, prompt_ind: int, cols=None):
        lo, hi = (0.0, bounds) if isinstance(bounds, float) else (bounds[0], bounds[1])
        a, b = int(lo * n_t), int(hi * n_t)
        sel = slice(None) if cols is None else torch.as_tensor(np.asarray(cols), dtype=torch.long)
        table[:, prompt_ind, sel] = 0
        table[a:b, prompt_ind, sel] = 1

    for i in range(len(prompts) - 1):
        paint(spec["default_"], i)
    for word, bounds in spec.items():
        if word == "default_":
            continue
        for i in range(1, len(prompts)):
            cols = get_word_inds(prompts[i], word, tokenizer)
            if len(cols) > 0:
                paint(bounds, i - 1, cols)
    return table.reshape(n_t, len(prompts) - 1, 1, 1, max_num_words)


def _align(x: Sequence[int], y: Sequence[int]) -> List[Tuple[int, int]]:
    """Needleman-Wunsch (gap 0, match +1, mismatch -1) with the reference's tie-breaking (left, then up, then diagonal);
    returns for every token of y the aligned token of x or -1."""
    nx, ny = len(x), len(y)
    score = np.zeros((nx + 1, ny + 1), dtype=np.int64)
    move = np.zeros((nx + 1, ny + 1), dtype=np.int8)  # 1 = left (gap in x), 2 = up (gap in y), 3 = diagonal
    move[0, 1:] = 1
    move[1:, 0] = 2
    for i in range(1, nx + 1):
        xi = x[i - 1]
        for j in range(1, ny + 1):
            cand = (score[i, j - 1], score[i - 1, j], score[i - 1, j - 1] + (1 if xi == y[j - 1] else -1))
            best = max(cand)
            score[i, j] = best
            move[i, j] = 1 if cand[0] == best else (2 if cand[1] == best else 3)
    out: List[Tuple[int, int]] = []
    i, j = nx, ny
    while i > 0 or j > 0:
        mv = move[i, j]
        if mv == 3:
            i, j = i - 1, j - 1
            out.append((j, i))
        elif mv == 1:
            j -= 1
            out.append((j, -1))
        else:
            i -= 1
    return out[::-1]


def get_refinement_mapper(prompts: Sequence[str], tokenizer, max_len: int = MAX_WORDS):
    """(mapper [P-1, 77] int64, alphas [P-1, 77]): target token n reads source token mapper[n] where alphas[n] == 1."""
    src = tokenizer.encode(prompts[0])
    mappers, alphas = [], []
    for tgt_prompt in prompts[1:]:
        tgt = tokenizer.encode(tgt_prompt)
        pairs = torch.tensor(_align(src, tgt), dtype=torch.int64)
        n = pairs.shape[0]
        al = torch.ones(max_len)
        al[:n] = (pairs[:, 1] != -1).float()
        mp = torch.zeros(max_len, dtype=torch.int64)
        mp[:n] = pairs[:, 1]
        mp[n:] = len(tgt) + torch.arange(max_len - len(tgt))
        mappers.append(mp)
        alphas.append(al)
    return torch.stack(mappers), torch.stack(alphas)


def get_replacement_mapper(prompts: Sequence[str], tokenizer, max_len: int = MAX_WORDS) -> torch.Tensor:
    """[P-1, 77, 77]: M[w, n] = weight of source token w in target token n (word-swap edits, equal word counts)."""
    out = []
    src_words = prompts[0].split(" ")
    for tgt_prompt in prompts[1:]:
        tgt_words = tgt_prompt.split(" ")
        if len(src_words) != len(tgt_words):
            raise ValueError("attention replacement edit can only be applied on prompts with the same length"
                             f" but prompt A has {len(src_words)} words and prompt B has {len(tgt_words)} words.")
        changed = [i for i, (a, b) in enumerate(zip(src_words, tgt_words)) if a != b]
        spans_src = [get_word_inds(prompts[0], i, tokenizer) for i in changed]
        spans_tgt = [get_word_inds(tgt_prompt, i, tokenizer) for i in changed]
        m = np.zeros((max_len, max_len))
        i = j = 0
        nxt = 0
        while i < max_len and j < max_len:
            if nxt < len(spans_src) and spans_src[nxt][0] == i:
                s_, t_ = spans_src[nxt], spans_tgt[nxt]
                if len(s_) == len(t_):
                    m[s_, t_] = 1
                else:
                    for col in t_:
                        m[s_, col] = 1.0 / len(t_)
                nxt += 1
                i += len(s_)
                j += len(t_)
            elif nxt < len(spans_src):
                m[i, j] = 1
                i, j = i + 1, j + 1
            else:
                m[j, j] = 1
                i, j = i + 1, j + 1
        out.append(torch.from_numpy(m).float())
    return torch.stack(out)


def get_equalizer(text: str, word_select, values, tokenizer=None) -> torch.Tensor:
    """[1, 77] multiplier per target token (Reweight)."""
    if isinstance(word_select, (int, str)):
        word_select = (word_select,)
    eq = torch.ones(1, MAX_WORDS)
    for word, val in zip(word_select, values):
        eq[:, get_word_inds(text, word, tokenizer)] = val
    return eq
