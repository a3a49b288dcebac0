"""FASTA ingest + token-budget batching (esme.fasta / esme.data drop-ins) against the reference's known
answers (its tests/test_data.py:11-32, tests/test_fasta.py) on its own fixture tests/golden/test.fa(.fai)."""
import pytest
import torch

from conftest import GOLDEN, load_golden
from esme.alphabet import tokenize_unpad
from esme.data import FastaDataset, FastaTokenDataset, TokenSizeBatchSampler
from esme.fasta import Fasta, build_fai, read_fai

FA, FAI = f'{GOLDEN}/test.fa', f'{GOLDEN}/test.fa.fai'
LENGTHS = [256, 320, 458, 156, 438, 60, 217, 204, 352, 75, 128, 447, 347, 948, 85, 137]


def test_read_fai_and_rebuild():
    fai = read_fai(FAI)
    assert len(fai) == 16 and [r['length'] for r in fai] == LENGTHS and set(fai[0]) == {
        'id', 'length', 'offset', 'line_bases', 'line_width'}
    assert build_fai(FA) == fai                       # the in-memory index equals samtools' file


def test_token_size_batch_sampler_known_answer():
    s = TokenSizeBatchSampler(LENGTHS, 400, shuffle=False)
    assert list(iter(s)) == [[0], [1], [2], [3], [4], [5, 6], [7], [8], [9, 10], [11], [12], [13], [14, 15]]
    assert len(s) == 13
    for seed in (0, 1):
        for batch in TokenSizeBatchSampler(LENGTHS, 1500, random_state=seed):
            assert sum(LENGTHS[i] + 2 for i in batch) <= 1500
    assert list(iter(TokenSizeBatchSampler(LENGTHS, 400, shuffle=False, drop_last=True)))[-1] == [13]


def test_fasta_reader():
    fa = Fasta(FA)
    assert len(fa) == 16 and [len(fa[i]) for i in range(16)] == LENGTHS
    assert fa[0].startswith('MAFSAEDVLKEYDRRRRMEALLLSLYYPNDRKLLDYKEWSPPRVQVECPKAPVEWNNP') and fa[0].endswith('GWKFTPL')
    assert fa['Q6GZX3'] == fa[1]
    assert len(Fasta(FA, max_len=100)) == 3
    with pytest.raises(FileNotFoundError):
        Fasta(f'{GOLDEN}/nope.fa')
    # the packed tokens of the whole file equal the fixture the reference produced from the same file
    g = load_golden('esm2_8m_testfa.npz')
    tok, idx, cu, ml = tokenize_unpad([fa[i] for i in range(16)])
    assert torch.equal(tok, g['tokens']) and torch.equal(cu, g['cu_lens']) and ml == g['max_len']


def test_fasta_token_dataset_batches():
    ds = FastaTokenDataset(FA, token_per_batch=1500, shuffle=False)
    seen = 0
    for tok, (cu, ml) in ds.to_dataloader():
        assert tok.dtype == torch.int64 and cu.dtype == torch.int32 and tok.numel() == int(cu[-1]) <= 1500
        assert ml == int((cu[1:] - cu[:-1]).max()) and isinstance(ml, int)
        seen += cu.numel() - 1
    assert seen == 16
    padded = FastaDataset(FA).to_dataloader(batch_size=4)
    assert next(iter(padded)).shape == (4, 460)


@pytest.mark.gpu
def test_device_batches_feed_the_model():
    import esme
    model = esme.ESM.from_pretrained(f'{GOLDEN}/esm2_8m.safetensors', device='cuda')
    ds = FastaTokenDataset(FA, token_per_batch=5000, shuffle=False, alphabet=esme.alphabet.Alphabet)
    outs = [model(t, pad) for t, pad in ds.device_batches('cuda')]
    assert len(outs) == 1 and outs[0].shape == (4660, 33)
    g = load_golden('esm2_8m_testfa.npz')
    rel = ((outs[0].float().cpu() - g['logits']).pow(2).mean().sqrt() / g['logits'].pow(2).mean().sqrt()).item()
    assert rel < 6e-3
