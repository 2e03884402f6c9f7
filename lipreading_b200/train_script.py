"""`train(**flags)` with the reference's signature, flag names and control flow
(src/scripts/train.py:134-354): seeded split by video, loaders without shuffling, initial eval,
epoch loop with patience -> lr/5 + restore-best, linear teacher-forcing decay, a fresh Adam every
epoch, best-CER checkpoints `best_encoder.pth` / `best_decoder.pth` under data/weights/<data>/<run#>.

Additions (all default-off so existing config files behave the same):
  frame_processing='flatten' | 'conv3d'   (VideoEncoder's own extension hook, better_model.py:10,22)
  max_epochs=-1                            hard stop for smoke runs
Under torchrun (WORLD_SIZE>1) training is data parallel (lipreading_b200/dist.py).
"""
import os
import time

import numpy as np
import torch

from . import data as _data_loader
from . import dist as _dist
from . import trainer as _train
from . import workspace as _util
from .model import CharDecodingStep, VideoEncoder

_log = _util.getLogger("train")


class _ScalarLog:
    """tensorboardX is not in this image; keep the same scalars in a CSV next to the weights."""

    def __init__(self, logdir):
        _util.mkdirP(logdir)
        self.path = os.path.join(logdir, "scalars.csv")

    def add_scalar(self, tag, value, global_step=None):
        with open(self.path, "a") as fh:
            fh.write("%s,%s,%s\n" % (tag, global_step, float(value)))

    def add_scalars(self, tag, values, global_step=None):
        for k, v in values.items():
            self.add_scalar(tag + "/" + k, v, global_step)


def _get_datasets(dataset_name, train_split, sentence_dataset, threshold=0.8, labels="labels.json", rand=None,
                  refresh=False, frame_type="face_lmk_seq"):
    ids = _data_loader.split_dataset(dataset_name, train_split=train_split, rand=rand)
    sets = [_data_loader.FrameCaptionDataset(dataset_name, split, vids, labels=labels, threshold=threshold,
                                             sentence_dataset=sentence_dataset, refresh=refresh,
                                             frame_type=frame_type)
            for split, vids in zip(("train", "val", "test"), ids)]
    print("\nDataset Information:")
    for name, ds in zip(("Train", "Val", "Test"), sets):
        print("\t%s Dataset Size: %d" % (name, len(ds)))
    print()
    return sets


def _init_models(char2idx, num_layers, frame_dim, hidden_size, char_dim, enable_ctc, rnn_type, attention_type,
                 attn_hidden_size, bidirectional, rnn_dropout, device, frame_processing="flatten"):
    encoder = VideoEncoder(frame_dim, hidden_size, frame_processing=frame_processing, rnn_type=rnn_type,
                           num_layers=num_layers, bidirectional=bidirectional, rnn_dropout=rnn_dropout,
                           enable_ctc=enable_ctc, vocab_size=len(char2idx), char2idx=char2idx,
                           device=device).to(device)
    decoding_step = CharDecodingStep(encoder, char_dim=char_dim, vocab_size=len(char2idx), char2idx=char2idx,
                                     rnn_dropout=rnn_dropout, attention_type=attention_type,
                                     attn_hidden_size=attn_hidden_size, device=device).to(device)
    return encoder, decoding_step


def restore(net, save_file):
    """Name- and shape-tolerant weight restore (src/scripts/train.py:82-132): copy what matches,
    report what does not, never raise on a missing / extra / reshaped variable."""
    own = net.state_dict()
    saved = torch.load(save_file, map_location="cpu")
    done = set()
    print("\tRestoring:")
    for name, value in saved.items():
        if name not in own:
            continue
        if own[name].size() != value.size():
            print("\t\tShape mismatch for var", name, "expected", own[name].size(), "got", value.size())
            continue
        own[name].copy_(value)
        done.add(name)
    ignored = sorted(set(saved) - done)
    unset = sorted(set(own) - done)
    print("\t\tRestored all variables" if not ignored else "\t\tDid not restore:\n\t" + "\n\t".join(ignored))
    print("\t\tNo new variables" if not unset else "\t\tInitialized but did not modify:\n\t" + "\n\t".join(unset))
    print("\tRestored %s" % save_file)


def train(
    data="StephenColbert/medium_no_vtx1",
    labels="labels.json",
    sentence_dataset=False,
    occlussion_threshold=0.8,
    train_split=0.8,
    num_workers=1,
    refresh=False,

    patience=10,
    batch_size=4,
    learning_rate=1e-4,
    annealings=2,
    enable_ctc=False,
    grad_norm=50,

    tr_epochs=50,
    max_tfr=0.9,
    min_tfr=0.0,

    num_layers=1,
    frame_dim=68 * 3,
    hidden_size=700,
    char_dim=300,

    rnn_type="LSTM",
    attention_type="1_layer_nn",
    attn_hidden_size=-1,
    bidirectional=False,
    rnn_dropout=0.0,

    seed=123456,
    cuda=False,

    frame_processing="flatten",
    max_epochs=-1,
):
    """ Runs the primary training loop (same flags as the reference's train.py; see module docstring). """
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    rand = np.random.RandomState(seed=seed)
    rank, local_rank, world = _dist.init()
    assert cuda or world == 1, "the kernels of this path run on CUDA only"
    assert cuda, "lipreading_b200 has no CPU path: pass --cuda"
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    print("Device: ", device)

    print("Initializing dataset '{}'".format(data))
    # the conv front-end eats the mouth-clip column (rows N2 -> N1), the reference's encoder the landmark column
    frame_type = "mouth_clip_seq" if frame_processing == "conv3d" else "face_lmk_seq"
    if world > 1 and refresh:
        # only rank 0 rebuilds the pickle cache; the others read it once it is complete
        if rank == 0:
            _get_datasets(data, train_split, sentence_dataset, threshold=occlussion_threshold, labels=labels,
                          rand=np.random.RandomState(seed=seed), refresh=True, frame_type=frame_type)
        _dist.barrier()
        refresh = False
    train_ds, val_ds, test_ds = _get_datasets(data, train_split, sentence_dataset, threshold=occlussion_threshold,
                                              labels=labels, rand=rand, refresh=refresh, frame_type=frame_type)

    def loader(ds):
        # no shuffle, no sampler (:209-211); batches are collated for the device (pad kernel / pinned prefetch) and
        # `batch_size` clips per GPU make a global batch of batch_size * world
        return _data_loader.GpuBatchLoader(ds, batch_size * world, device, rank=rank, world=world)
    train_loader, val_loader, test_loader = loader(train_ds), loader(val_ds), loader(test_ds)
    reducer = _dist.GradAllReducer(world) if world > 1 else None

    print("Initializing model")
    encoder, decoding_step = _init_models(train_ds.char2idx, num_layers, frame_dim, hidden_size, char_dim,
                                          enable_ctc, rnn_type, attention_type, attn_hidden_size, bidirectional,
                                          rnn_dropout, device, frame_processing)
    # one run directory for the whole job: rank 0 picks the first free index, the others are told
    weights_dir = _dist.broadcast_object(_util.getRelWeightsPath(data, use_existing=False) if rank == 0 else None)
    if rank == 0:
        _util.mkdirP(weights_dir)
    _dist.barrier()
    writer = _ScalarLog(weights_dir) if rank == 0 else None
    encoder_path = os.path.join(weights_dir, "best_encoder.pth")
    decoder_path = os.path.join(weights_dir, "best_decoder.pth")

    def cer_of(loader_):
        # every rank evaluates its shard; the hit / character counts are summed over the ranks BEFORE the ratio, so
        # val_cer — which drives the loop condition, the annealing and the checkpoints — is identical everywhere
        _, correct, count = _train.eval(encoder, decoding_step, loader_, device, train_ds.char2idx)
        correct, count = _dist.allreduce_counts(correct, count, device=device)
        return ((count - correct) / count).float()

    print("Initial evaluation...")
    val_cer = cer_of(val_loader)
    print("\tCER: ", str(val_cer))

    val_cers, dec_losses, ctc_losses = [], [], []
    best_val_cer, best_idx = 1.0, -1
    num_epochs, num_annealings = 0, 0
    print("Beginning training loop")
    ts = time.time()
    while val_cer < best_val_cer or num_annealings < annealings:
        if 0 <= max_epochs <= num_epochs:
            break
        print("Epoch {}:".format(num_epochs + 1))
        if num_epochs - best_idx > patience:
            num_annealings += 1
            learning_rate /= 5
            print(f"\tAnnealing to {learning_rate}")
            _dist.barrier()                          # rank 0 has finished writing the checkpoints
            if os.path.isfile(encoder_path):
                restore(encoder, encoder_path)
                restore(decoding_step, decoder_path)
            best_idx = num_epochs
        curr_tfr = max(min_tfr, max_tfr - num_epochs / tr_epochs)
        assert 0.0 <= curr_tfr <= 1.0
        print(f"\tCurrent Teacher Forcing Ratio: {curr_tfr}")
        opt = torch.optim.Adam(list(encoder.parameters()) + list(decoding_step.parameters()), lr=learning_rate)
        avg_dec, avg_ctc = _train.train(encoder, decoding_step, train_loader, opt=opt, device=device,
                                        char2idx=train_ds.char2idx, teacher_forcing_ratio=curr_tfr,
                                        grad_norm=grad_norm, dist=reducer)
        print(f"\tAVG Decoder Loss: {avg_dec}\n\tAVG CTC Loss: {avg_ctc}")
        val_cer = cer_of(val_loader)
        train_cer = cer_of(train_loader)
        if rank == 0:
            encoder.save_best_model(val_cer, encoder_path)
            decoding_step.save_best_model(val_cer, decoder_path)
            writer.add_scalar(os.path.join(data, "avg decoder loss"), avg_dec, global_step=num_epochs)
            writer.add_scalar(os.path.join(data, "avg CTC loss"), avg_ctc, global_step=num_epochs)
            writer.add_scalars(os.path.join(data, "CER"), {"Train": train_cer, "Val": val_cer}, global_step=num_epochs)
            writer.add_scalar(os.path.join(data, "learning rate"), learning_rate, global_step=num_epochs)
        else:
            # keep best_error in step with rank 0 (it gates the next save there and nothing here)
            encoder.best_error = min(encoder.best_error, float(val_cer))
            decoding_step.best_error = min(decoding_step.best_error, float(val_cer))
        print(f"\tTrain CER: {train_cer}\n\tVal CER: {val_cer}")
        with torch.no_grad():
            print(f"\tTest CER: {cer_of(test_loader)}")
        val_cers.append(float(val_cer))
        dec_losses.append(avg_dec)
        ctc_losses.append(avg_ctc)
        if val_cer < best_val_cer:
            best_val_cer, best_idx = val_cer, num_epochs
        num_epochs += 1

    total = time.time() - ts
    print("\nTraining complete: Took '{}' seconds, or '{}' per epoch".format(total, total / max(num_epochs, 1)))
    if val_cers:
        print("Training Statistics\n\tBest Val CER: '{}'\n\tBest Decoder Loss: '{}'\n\tBest CTC Loss: '{}'\n".format(
            np.min(val_cers), np.min(dec_losses), np.min(ctc_losses)))
    return {"val_cers": val_cers, "dec_losses": dec_losses, "ctc_losses": ctc_losses, "weights_dir": weights_dir}


def main(argv=None):
    from .cli import parseArgsForClassOrScript
    args = vars(parseArgsForClassOrScript(train, argv))
    args.pop("verbosity", None)
    train(**args)


if __name__ == "__main__":
    main()
