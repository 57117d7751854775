"""MIDI file -> wav on the B200 path: the argument list of the reference's synthesize_midi_file.py
(:12-36) over load_midi_as_conditioning -> PianoModel -> 16-bit wav.  A thin caller of the hot path
(DESIGN.md section 8), not a port of the reference's CLI: gin files are replaced by the two model
factories (--model v2 | dafx22) and the checkpoint is read without TensorFlow.

usage: python scripts/synthesize_midi_file.py [--model v2] [--ckpt PREFIX_OR_NPZ] [--piano_type 9]
           [-wu 0.5] [-d SECONDS] [-n DBFS] [-u] midi_file out_file
"""
import argparse
import os
import sys
import wave

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

MODELS = {'v2': ('maestro_v2_model', 'tests/golden/v2_weights.npz', 24000),
          'dafx22': ('dafx22_model', 'tests/golden/dafx22_weights.npz', 16000)}


def process_args(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('-m', '--model', choices=sorted(MODELS), default='v2',
                        help='maestro-v2.gin (24 kHz, the reference default) or dafx22.gin (16 kHz)')
    parser.add_argument('--ckpt', type=str, default=None,
                        help='TensorFlow checkpoint prefix (e.g. model_weights/v2/ckpt-225000) or .npz fixture')
    parser.add_argument('--piano_type', type=int, default=9, help='Piano model (from 0 to 9)')
    parser.add_argument('-wu', '--warm_up', type=float, default=0.5, help='Warm-up duration (in s)')
    parser.add_argument('-d', '--duration', type=float, default=None, help='Maximum duration of synthesized audio')
    parser.add_argument('-n', '--normalize', type=float, default=None, help='Normalize audio to this amount of dBFS')
    parser.add_argument('-u', '--unreverbed', action='store_true', help='Also write the dry audio')
    parser.add_argument('midi_file', type=str)
    parser.add_argument('out_file', type=str)
    return parser.parse_args(argv)


def normalize_dbfs(audio, volume):
    """io_utils.py:245-253 (pydub: dBFS of the RMS against full scale, one gain for the file)."""
    rms = float(np.sqrt(np.mean(audio.astype(np.float64) ** 2)))
    if rms == 0.0:
        return audio
    return audio * np.float32(10.0 ** ((volume - 20.0 * np.log10(rms)) / 20.0))


def write_wav(path, audio, sample_rate):
    """Mono 16-bit PCM (what soundfile.write picks for a .wav file by default); out-of-range samples clip."""
    pcm = np.clip(np.rint(audio.astype(np.float64) * 32768.0), -32768, 32767).astype('<i2')
    with wave.open(path, 'wb') as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(int(sample_rate))
        f.writeframes(pcm.tobytes())


def main(args):
    import ddsp_piano_b200 as dp
    from ddsp_piano_b200 import midi
    factory, default_ckpt, sample_rate = MODELS[args.model]
    inputs = midi.load_midi_as_conditioning(args.midi_file, duration=args.duration,
                                            warm_up_duration=args.warm_up)          # :42-44
    inputs['piano_model'] = np.array([[args.piano_type]], np.int64)                 # :46
    model = getattr(dp, factory)(args.ckpt or os.path.join(ROOT, default_ckpt), device='cuda:0',
                                 sample_rate=sample_rate, inference=True)
    outs = model(inputs)                                                            # :74
    skip = int(args.warm_up * sample_rate)
    todo = [(args.out_file, outs['audio_synth'])]                                   # :77-79
    if args.unreverbed:
        todo.append((args.out_file + '_unreverbed.wav', outs['add']['signal']))     # :84-87
    for path, signal in todo:
        audio = signal[0, skip:].cpu().numpy()
        if args.normalize:
            audio = normalize_dbfs(audio, args.normalize)
        write_wav(path, audio, sample_rate)
    print(f"{inputs['duration'] - args.warm_up:.2f} s of audio at {sample_rate} Hz saved at {args.out_file}")


if __name__ == '__main__':
    main(process_args())
