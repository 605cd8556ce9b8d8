#include "FastqChunker.h"

#include <dirent.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

FastqChunker::FastqChunker(const std::string &directory) {
    if (DIR *dir = opendir(directory.c_str())) {
        while (struct dirent *ent = readdir(dir)) {
            if (ent->d_name[0] == '.') continue;                 // InputFileHandler.cpp:30
            _files.push_back(directory + "/" + ent->d_name);
        }
        closedir(dir);
    }
    std::sort(_files.begin(), _files.end());                     // the artefact is order-independent; be deterministic
    if (!_files.empty()) {
        if (FILE *f = fopen(_files[0].c_str(), "rb")) {          // L := length of line 2 (FASTQFileReader.cpp:30-35)
            char *buf = nullptr;
            size_t cap = 0;
            for (int i = 0; i < 2; i++) {
                ssize_t n = getline(&buf, &cap, f);
                if (n < 0) { n = 0; }
                while (n > 0 && (buf[n - 1] == '\n' || buf[n - 1] == '\r')) n--;
                if (i == 1) _lineLength = n;
            }
            free(buf);
            fclose(f);
        }
    }
}

FastqChunker::~FastqChunker() {
    if (_fp) fclose(_fp);
    free(_lineBuf);
}

bool FastqChunker::openNext() {
    if (_fp) { fclose(_fp); _fp = nullptr; }
    while (_fileIndex < _files.size()) {
        _fp = fopen(_files[_fileIndex++].c_str(), "rb");
        _prev.clear();
        if (_fp) return true;
    }
    return false;
}

bool FastqChunker::nextLine(std::string &out) {
    if (!_fp) return false;
    ssize_t n = getline(&_lineBuf, &_lineCap, _fp);
    if (n < 0) return false;
    while (n > 0 && (_lineBuf[n - 1] == '\n' || _lineBuf[n - 1] == '\r')) n--;
    out.assign(_lineBuf, (size_t)n);
    return true;
}

int64_t FastqChunker::read(char *dst, int64_t capacity) {
    int64_t off = 0;
    const int64_t L = _lineLength;
    if (L <= 0) return 0;
    if (_havePending) {                                          // a read that did not fit the previous chunk
        if (L > capacity) return 0;
        memcpy(dst, _pending.data(), (size_t)L);
        off = L;
        _havePending = false;
    }
    std::string line;
    while (off + L <= capacity) {
        if (!_fp && !openNext()) break;
        if (!nextLine(line)) {                                   // end of this file
            fclose(_fp);
            _fp = nullptr;
            continue;
        }
        if (!line.empty() && line[0] == '+') {                   // the line before it is a sequence (:57-79)
            if ((int64_t)_prev.size() == L) {
                memcpy(dst + off, _prev.data(), (size_t)L);
                off += L;
                _reads++;
            } else if (!_prev.empty()) {
                _skipped++;
            }
            nextLine(line);                                      // the quality line is never inspected
            _prev.clear();
        } else {
            _prev.swap(line);
        }
    }
    return off;
}
