"""The training driver of the reference (``src/models/train_model.py:772-1048``) over the CUDA path: same flags, same data-set
format, same iterator semantics, same per-epoch statistics and checkpoint files.

    python -m pivp_b200.train_loop --data_dir=data/processed/.../push_train --batch_size=32 --gpu=0 ...

What is kept to the letter: the 19 flag names and defaults (:773-791), the output directory ``YYYYmmdd-HHMMSS-<model_type>-<batch>``
(:806; predict_model.py:92-95 parses the model type from it), ``training-<epoch>`` / ``state-<epoch>`` npz files and the
``training-global_*`` statistics (:1035-1041), the ``version`` file with the git branch and commit (:874-885, 1031-1033), the
``[mean, std, min, max, median]`` epoch rows (:970-973), the SerialIterator epoch logic (:914, 937-940).
What differs on purpose: the batch goes to the device through the double-buffered ``BatchPrefetcher`` instead of a synchronous
``xp.array`` (:950); the step is one replayed CUDA graph; the reference's validation never runs (``epoch+1 % validation_interval``
is never zero, :981; ``xp.act_validation_set`` at :992 would raise) -- here it runs every ``validation_interval`` epochs in test mode;
``--pretrained_state`` is loaded into the optimizer (the reference loads it into the model, :868).
"""
import logging
import os
import subprocess
import time

import numpy as np


def git_version():
    """train_model.py:874-885: '<branch>\\n<commit>' of the working tree, or None outside a git checkout."""
    try:
        run = lambda *a: subprocess.check_output(("git",) + a, stderr=subprocess.DEVNULL).decode().strip()
        return run("rev-parse", "--abbrev-ref", "HEAD") + "\n" + run("rev-parse", "HEAD")
    except Exception:
        return None


def epoch_row(values):
    """train_model.py:970-973."""
    v = np.asarray(values, np.float64)
    return [v.mean(), v.std(), v.min(), v.max(), np.median(v)]


def train(data_dir='data/processed/brain-robotics-data/push/push_train', output_dir='models', event_log_dir='models', num_iterations=100000,
          pretrained_model='', pretrained_state='', sequence_length=10, context_frames=2, use_state=1, model_type='CDNA', num_masks=10,
          schedsamp_k=900.0, train_val_split=0.95, batch_size=32, learning_rate=0.001, gpu=0, validation_interval=200, save_interval=50,
          debug=0, compute="bf16", graph=True, log_every=1):
    """``main`` of train_model.py (:792-1048).  Returns (global_losses, global_psnr_all, save_dir)."""
    import torch
    from . import data as D
    from .links import Model, Adam, concat_examples
    from .train import TrainStep, BatchPrefetcher
    from .serializers import save_npz, load_npz
    logger = logging.getLogger(__name__)
    logger.info('Training the model')
    logger.info('Model: {}'.format(model_type))
    logger.info('GPU: {}'.format(gpu))
    logger.info('# Minibatch-size: {}'.format(batch_size))
    logger.info('# Num iterations: {}'.format(num_iterations))
    logger.info('# epoch: {}'.format(round(num_iterations / batch_size)))
    model_suffix_dir = "{0}-{1}-{2}".format(time.strftime("%Y%m%d-%H%M%S"), model_type, batch_size)          # :806
    images, actions, states = D.load_dataset(data_dir)                                                       # :812-834
    grouped_training, grouped_validation = D.train_val_split(images, actions, states, train_val_split)      # :836-843, 896-911
    logger.info('Data set contain {0}, {1} will be use for training and {2} will be use for validation'.format(
        len(images), len(grouped_training), len(grouped_validation)))
    T, H, W = images.shape[1], images.shape[2], images.shape[3]
    dev = "cuda:%d" % max(gpu, 0)                              # this path has no CPU fallback: --gpu=-1 (the reference's CPU mode) runs on device 0
    model = Model(num_masks=num_masks, is_cdna=model_type == 'CDNA', is_dna=model_type == 'DNA', is_stp=model_type == 'STP',
                  use_state=bool(use_state), scheduled_sampling_k=schedsamp_k, num_frame_before_prediction=context_frames, prefix='train',
                  height=H, width=W, device=dev, compute=compute)                                            # :848-857
    optimizer = Adam(alpha=learning_rate).setup(model)                                                       # :860-861
    if pretrained_model:
        load_npz(pretrained_model, model)
        logger.info("Loading pretrained model {}".format(pretrained_model))
    if pretrained_state:
        load_npz(pretrained_state, optimizer)
        logger.info("Loading pretrained state {}".format(pretrained_state))
    current_version = git_version()
    train_iter = D.SerialIterator(grouped_training, batch_size, repeat=True, shuffle=True)                   # :914
    valid_iter = D.SerialIterator(grouped_validation, batch_size, repeat=False, shuffle=True)
    step = TrainStep(model, optimizer, batch_size, T, graph=graph)
    pin = lambda arrs: [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in arrs]

    epochs_of = []                                             # (epoch, is_new_epoch) of every batch handed to the prefetcher, in order

    def batches():
        for _ in range(num_iterations):
            ep = train_iter.epoch
            b = train_iter.next()
            epochs_of.append((ep, train_iter.is_new_epoch, train_iter.current_position))
            yield pin(concat_examples(b))
    pf = BatchPrefetcher(step, batches())

    local_losses, local_psnr_all = [], []
    global_losses, global_psnr_all, global_losses_valid, global_psnr_all_valid = [], [], [], []
    save_dir = output_dir + '/' + model_suffix_dir
    start_time, itr = None, 0
    while itr < num_iterations and pf.load_next(prefetch=False):                                             # :937
        epoch, is_new_epoch, pos = epochs_of[itr]
        if start_time is None:
            start_time = time.time()
        loss = step(itr)                                                                                     # :950 optimizer.update(...)
        pf.prefetch()                                          # draw batch itr+1 AFTER this step's scheduled-sampling shuffles (same global RNG order as :939-950)
        local_losses.append(float(loss))                                                                     # :955-956 (D2H of two scalars)
        local_psnr_all.append(float(model.psnr_all))
        model.reset_state()                                                                                  # :960
        if itr % log_every == 0:
            logger.info("Global iteration: {}  epoch {}  mini-batch {}/{}  loss {}".format(itr + 1, epoch + 1, pos, len(grouped_training), local_losses[-1]))
        if is_new_epoch:                                                                                     # :964-979
            logger.info("[TRAIN] Epoch #: {}".format(epoch + 1))
            logger.info("[TRAIN] Epoch elapsed time: {}".format(time.time() - start_time))
            global_losses.append(epoch_row(local_losses))
            global_psnr_all.append(epoch_row(local_psnr_all))
            logger.info("[TRAIN] epoch loss: {}".format(global_losses[-1][0]))
            logger.info("[TRAIN] epoch psnr: {}".format(global_psnr_all[-1][0]))
            local_losses, local_psnr_all = [], []
            start_time = None
            optimizer.new_epoch()
        if is_new_epoch and (epoch + 1) % validation_interval == 0 and len(grouped_validation) >= batch_size:   # :981-1021, fixed
            vl, vp = [], []
            model.train = False                                # chainer.using_config('train', False)
            try:
                for vb in valid_iter:
                    if len(vb) != batch_size:
                        break
                    model([torch.from_numpy(a) for a in concat_examples(vb)], itr)
                    vl.append(float(model.loss)); vp.append(float(model.psnr_all))
                    model.reset_state()
            finally:
                model.train = True
            if vl:
                global_losses_valid.append(epoch_row(vl)); global_psnr_all_valid.append(epoch_row(vp))
                logger.info("[VALID] epoch loss: {}".format(global_losses_valid[-1][0]))
                logger.info("[VALID] epoch psnr: {}".format(global_psnr_all_valid[-1][0]))
            valid_iter.reset()
        if is_new_epoch and epoch % save_interval == 0:                                                      # :1023-1041
            logger.info('Saving model')
            if not os.path.exists(save_dir):
                os.makedirs(save_dir)
                with open(save_dir + '/version', 'w') as f:
                    f.write(str(current_version) + '\n')
            save_npz(save_dir + '/training-' + str(epoch), model)
            save_npz(save_dir + '/state-' + str(epoch), optimizer)
            np.save(save_dir + '/training-global_losses', np.array(global_losses))
            np.save(save_dir + '/training-global_psnr_all', np.array(global_psnr_all))
            np.save(save_dir + '/training-global_losses_valid', np.array(global_losses_valid))
            np.save(save_dir + '/training-global_psnr_all_valid', np.array(global_psnr_all_valid))
        itr += 1
    return global_losses, global_psnr_all, save_dir


FLAGS = [  # (name, type, default, help) -- train_model.py:773-791
    ('data_dir', str, 'data/processed/brain-robotics-data/push/push_train', 'Directory containing data.'),
    ('output_dir', str, 'models', 'Directory for model checkpoints.'),
    ('event_log_dir', str, 'models', 'Directory for writing summary.'),
    ('num_iterations', int, 100000, 'Number of training iterations. Number of epoch is: num_iterations/batch_size.'),
    ('pretrained_model', str, '', 'Filepath of a pretrained model to initialize from.'),
    ('pretrained_state', str, '', 'Filepath of a pretrained state to initialize from.'),
    ('sequence_length', int, 10, 'Sequence length, including context frames.'),
    ('context_frames', int, 2, 'Number of frames before predictions.'),
    ('use_state', int, 1, 'Whether or not to give the state+action to the model.'),
    ('model_type', str, 'CDNA', 'Model architecture to use - CDNA, DNA, or STP.'),
    ('num_masks', int, 10, 'Number of masks, usually 1 for DNA, 10 for CDNA, STP.'),
    ('schedsamp_k', float, 900.0, 'The k parameter for schedules sampling. -1 for no scheduled sampling.'),
    ('train_val_split', float, 0.95, 'The percentage of data to use for the training set, vs. the validation set.'),
    ('batch_size', int, 32, 'Batch size for training.'),
    ('learning_rate', float, 0.001, 'The base learning rate of the generator.'),
    ('gpu', int, -1, 'ID of the gpu(s) to use'),
    ('validation_interval', int, 200, 'How often to run a batch through the validation model'),
    ('save_interval', int, 50, 'How often to save a model checkpoint'),
    ('debug', int, 0, 'Debug mode.'),
]


def main(argv=None):
    import argparse
    ap = argparse.ArgumentParser(description="Train the model based on the data saved in ../processed (train_model.py:main)")
    for name, typ, default, hlp in FLAGS:
        ap.add_argument('--' + name, type=typ, default=default, help=hlp)
    ap.add_argument('--compute', default='bf16', choices=['bf16', 'f32'], help='tensor-core bf16 path (default) or the fp32 parity path')
    args = ap.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format='%(asctime)s - %(name)s - %(levelname)s - %(message)s')       # :1052-1055
    return train(**vars(args))


if __name__ == '__main__':
    main()
