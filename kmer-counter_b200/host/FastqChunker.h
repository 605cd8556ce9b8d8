// FastqChunker.h -- host-side reader that produces the seam's input layout.
//
// Mirrors what InputFileHandler + FASTQFileReader hand to processKMers
// (InputFileHandler.cpp:22-47,82-95; FASTQFileReader.cpp:18-89): every non-dot file of a
// directory, read length L := length of line 2 of the first file, a sequence is "the
// line before a line that starts with '+'", sequences are copied back to back (no
// separators) into a chunk.  Differences, all deliberate:
//   * a chunk is filled to capacity (the reference stops one read early);
//   * sequences whose length differs from L are skipped and counted instead of silently
//     breaking the fixed stride;
//   * a file without a trailing newline does not re-emit its last quality line.
#pragma once

#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

class FastqChunker {
public:
    explicit FastqChunker(const std::string &directory);
    ~FastqChunker();
    bool ok() const { return !_files.empty() && _lineLength > 0; }
    int64_t getLineLength() const { return _lineLength; }
    // Fills dst (capacity bytes) with whole reads; returns bytes written, 0 at end of input.
    int64_t read(char *dst, int64_t capacity);
    const std::vector<std::string> &files() const { return _files; }
    uint64_t skippedReads() const { return _skipped; }
    uint64_t totalReads() const { return _reads; }

private:
    bool nextLine(std::string &out);
    bool openNext();
    std::vector<std::string> _files;
    size_t _fileIndex = 0;
    FILE *_fp = nullptr;
    char *_lineBuf = nullptr;
    size_t _lineCap = 0;
    int64_t _lineLength = 0;
    std::string _prev, _pending;
    bool _havePending = false;
    uint64_t _skipped = 0, _reads = 0;
};
