// oracle/shim: boost::iostreams::filtering_ostream as used by logger.cpp:78-155 --
// a chain of optional tee-to-ostream filters ending in a file sink or a null sink.
#pragma once
#include "../config.hpp"
#include <fstream>
#include <memory>
#include <ostream>
#include <streambuf>
namespace boost {
namespace iostreams {
template <class Stream> struct tee_filter {
  explicit tee_filter(Stream &s) : target(&s) {}
  Stream *target;
};
struct file_sink {
  explicit file_sink(const std::string &path)
      : file(std::make_shared<std::ofstream>(path.c_str())) {}
  bool is_open() const { return file->is_open(); }
  std::shared_ptr<std::ofstream> file;
};
struct null_sink {};
class filtering_ostream : public std::ostream {
  class chainbuf : public std::streambuf {
  public:
    std::vector<std::ostream *> tees;
    std::shared_ptr<std::ofstream> file;
  protected:
    int_type overflow(int_type ch) override {
      if (ch != traits_type::eof()) {
        char c = static_cast<char>(ch);
        for (auto *t : tees) t->put(c);
        if (file) file->put(c);
      }
      return ch;
    }
    std::streamsize xsputn(const char *s, std::streamsize n) override {
      for (auto *t : tees) t->write(s, n);
      if (file) file->write(s, n);
      return n;
    }
    int sync() override {
      for (auto *t : tees) t->flush();
      if (file) file->flush();
      return 0;
    }
  };
  chainbuf buf;
public:
  filtering_ostream() : std::ostream(nullptr) { rdbuf(&buf); }
  template <class S> void push(const tee_filter<S> &t) { buf.tees.push_back(t.target); }
  void push(const file_sink &f) { buf.file = f.file; }
  void push(const null_sink &) {}
  void reset() {
    flush();
    buf.tees.clear();
    buf.file.reset();
  }
};
} // namespace iostreams
} // namespace boost
