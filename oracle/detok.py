"""TEST INFRASTRUCTURE -- CPU restatement of the reference's token -> MIDI-note loop (api_cache.py:157,208-221).

Only tests/ may import this.  The reference feeds pretty_midi objects (third-party, not installed here); this restatement
returns the same information as plain data: the instruments in order of appearance with the notes each one received.
The two pretty_midi look-ups are parameters (``program_of``, ``number_of``) so that the same functions drive the oracle and the
table builder of the engine (vocab.note_table).  Parity unpinned against pretty_midi itself (absent); pinned against the
reference's control flow by construction: every branch below cites its line.
"""
import re

NOTE_RE = re.compile(r"\[NOTE\] \[PITCH:(.+?)\] \[START:(.+?)\] \[END:(.+?)\] \[DURATION:(.+?)\]")   # api_cache.py:157


def tokens_to_instruments(tokens, program_of, number_of):
    instruments, current = [], None                                   # api_cache.py:208
    for tok in tokens:                                                 # :209
        if tok.startswith("[INSTRUMENT]"):                             # :210
            name = tok.split("]", 1)[1].strip()                        # :211
            current = {"name": name, "program": program_of(name), "notes": []}   # :212-214
            instruments.append(current)
        else:
            m = NOTE_RE.match(tok)                                     # :215
            if m and current is not None:
                pitch = number_of(m.group(1))                          # :216
                start, end = float(m.group(2)), float(m.group(3))      # :217
                current["notes"].append((pitch, start, end))           # :218-220 (velocity is the constant 100)
    return instruments
