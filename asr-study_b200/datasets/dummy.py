"""Synthetic dataset with the semantics of datasets/dummy.py:12-112 ("fake dataset reader and
parser to do some tests"), seeded and kept in memory instead of temp wav files + librosa:
num_speakers x num_utterances_per_speaker clips, duration ~ U(min,max) s of Gaussian noise,
labels = random chars in a..y of length randint(2, max_label_length), optional
split=[train, valid] fractions -> 'train' / 'valid' / 'test'."""
import numpy as np


class Dummy(object):
    def __init__(self, num_speakers=10, num_utterances_per_speaker=10, max_duration=10.0, min_duration=1.0,
                 max_label_length=50, fs=16e3, split=None, seed=1234, name="dummy"):
        if split is not None and (len(split) != 2 or np.sum(split) > 1.):
            raise ValueError("Split must have len = 2 and must sum <= 1")
        self.num_speakers, self.num_utt = num_speakers, num_utterances_per_speaker
        self.max_duration, self.min_duration = max_duration, min_duration
        self.max_label_length, self.fs, self.split, self.seed, self.name = max_label_length, fs, split, seed, name

    def __iter__(self):
        rng = np.random.RandomState(self.seed)
        total = self.num_speakers * self.num_utt
        counter = 0
        for speaker in range(self.num_speakers):
            for _ in range(self.num_utt):
                duration = rng.uniform(low=self.min_duration, high=self.max_duration)
                audio = rng.randn(int(np.floor(duration * self.fs))).astype(np.float32)
                label = rng.randint(low=ord("a"), high=ord("z"), size=(rng.randint(2, self.max_label_length),))
                data = {"duration": duration, "input": audio, "label": "".join(chr(c) for c in label),
                        "speaker": "speaker_%d" % speaker}
                if self.split is not None:
                    if counter < np.floor(self.split[0] * total):
                        data["dataset"] = "train"
                    elif counter < np.floor(np.sum(self.split) * total):
                        data["dataset"] = "valid"
                    else:
                        data["dataset"] = "test"
                counter += 1
                yield data

    def to_dict_list(self):
        rows = list(self)
        return {k: [r[k] for r in rows] for k in rows[0]}
