// Drives the C++ shim exactly like the reference's call sites do (Frame::createStereo / ORBExtractor / createRGBD) on raw
// image files written by tests/test_cpp_shim.py, and dumps the results for comparison with the oracle.
//   shim_main stereo <w> <h> <nFeatures> <nLevels> <scale> <template> <left.raw> <right.raw> <out.bin>
//   shim_main rgbd   <w> <h> <nFeatures> <nLevels> <scale> <template> <gray.raw> <depth_u16.raw> <out.bin> <dScale>
//   shim_main errors <template>
//   shim_main threads <w> <h> <nFeatures> <nLevels> <scale> <template> <a.raw> <b.raw> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <thread>

#include <orbx/orb_slam2_shim.hpp>

using namespace ORB_SLAM2_ROS2_B200;

static cv::Mat read_raw(const char *path, int w, int h, int type)
{
  cv::Mat m(h, w, type);
  std::ifstream f(path, std::ios::binary);
  f.read((char *)m.data, (std::streamsize)((size_t)w * h * m.elemSize()));
  if (!f) throw std::runtime_error(std::string("cannot read ") + path);
  return m;
}

template <typename T> static void put(std::ofstream &o, const T &v) { o.write((const char *)&v, sizeof(T)); }

static void dump(std::ofstream &o, const std::vector<cv::KeyPoint> &k, const std::vector<cv::Mat> &d)
{
  put(o, (int32_t)k.size());
  for (auto &kp : k) o.write((const char *)&kp, sizeof(cv::KeyPoint));
  for (auto &m : d) o.write((const char *)m.data, 32);
}

int main(int argc, char **argv)
{
  try
  {
    std::string mode = argv[1];
    if (mode == "errors")
    {
      int ok = 0;
      cv::Mat img(240, 320, CV_8U);
      try
      {
        ORBExtractor ex(img, 500, 4, 1.2f, "/nonexistent/brief_template.txt", 20, 7);
      }
      catch (const FileNotOpenError &)
      {
        ++ok;
      }
      cv::Mat tiny(60, 100, CV_8U);
      try
      {
        ORBExtractor ex(tiny, 500, 8, 1.2f, argv[2], 20, 7);
      }
      catch (const ImageSizeError &)
      {
        ++ok;
      }
      std::printf("errors ok=%d\n", ok);
      return ok == 2 ? 0 : 1;
    }
    int w = atoi(argv[2]), h = atoi(argv[3]), nf = atoi(argv[4]), nl = atoi(argv[5]);
    float scale = (float)atof(argv[6]);
    std::string tmpl = argv[7];
    std::ofstream out(argv[10], std::ios::binary);
    if (mode == "stereo")
    {
      Camera::set(718.856f, 718.856f, 607.1928f, 185.2157f, 0.537166f);
      cv::Mat l = read_raw(argv[8], w, h, CV_8U), r = read_raw(argv[9], w, h, CV_8U);
      // the extractor on its own, then the frame factory -- as Tracking::grabFrame would (src/Tracking.cc:82-88)
      ORBExtractor ex(l, nf, nl, scale, tmpl, 20, 7);
      std::vector<cv::KeyPoint> k;
      std::vector<cv::Mat> d;
      ex.extract(k, d);
      dump(out, k, d);
      put(out, (int32_t)ex.getPyramid().size());
      put(out, (int32_t)ex.getPyramid().back().cols);
      put(out, ORBExtractor::getScaledFactors().back());
      Frame::SharedPtr f = Frame::createStereo(l, r, nf, tmpl, 20, 7, nullptr, nl, scale);
      dump(out, f->getLeftKeyPoints(), f->getLeftDescriptor());
      dump(out, f->getRightKeyPoints(), f->getRightDescriptor());
      put(out, (int32_t)f->getN());
      for (double v : f->getRightU()) put(out, v);
      for (double v : f->getDepth()) put(out, v);
      // VirtualFrame::initGrid: every keypoint sits in exactly one cell, cells are ascending
      auto grids = f->getGrids();
      std::size_t total = 0;
      bool ascending = true;
      for (auto &row : grids)
        for (auto &cell : row)
        {
          total += cell.size();
          for (std::size_t i = 1; i < cell.size(); ++i) ascending = ascending && cell[i - 1] < cell[i];
        }
      put(out, (int32_t)grids.size());
      put(out, (int32_t)grids[0].size());
      put(out, (int32_t)total);
      put(out, (int32_t)ascending);
      // constant-velocity matching of the frame against itself shifted by (+2, -1) px, every 3rd keypoint of "frame 1" taken
      std::vector<cv::KeyPoint> k2 = f->getLeftKeyPoints();
      for (auto &kp : k2) kp.pt.x += 2.f, kp.pt.y -= 1.f;
      std::vector<bool> valid2(k2.size(), true), hasMp1(k2.size(), false);
      for (std::size_t i = 0; i < hasMp1.size(); i += 3) hasMp1[i] = true;
      ORBMatcher matcher(0.6f);
      std::vector<cv::DMatch> matches;
      int nm = matcher.searchByProjection(f, k2, f->getLeftDescriptor(), valid2, hasMp1, matches, 15.f);
      put(out, (int32_t)nm);
      for (auto &m : matches) put(out, (int32_t)m.queryIdx), put(out, (int32_t)m.trainIdx), put(out, m.distance);
      ORBMatcher::verifyAngle(f, matches, f->getLeftKeyPoints(), k2);
      put(out, (int32_t)matches.size());
      for (auto &m : matches) put(out, (int32_t)m.queryIdx), put(out, (int32_t)m.trainIdx), put(out, m.distance);
      // KeyFrame::serializeToProtobuf for a keyframe made from this frame
      std::string rec = f->serializeKeyFrameData(42);
      put(out, (int32_t)rec.size());
      out.write(rec.data(), (std::streamsize)rec.size());
      // VirtualFrame::computeBow with a text vocabulary (optional 11th argument)
      if (argc > 11)
      {
        Vocabulary voc(argv[11], f->context());
        BowVector bv;
        FeatureVector fv;
        f->computeBow(voc, bv, fv);
        put(out, (int32_t)voc.size());
        put(out, (int32_t)bv.size());
        for (auto &kv : bv) put(out, (int32_t)kv.first), put(out, kv.second);
        put(out, (int32_t)fv.size());
        for (auto &kv : fv)
        {
          put(out, (int32_t)kv.first);
          put(out, (int32_t)kv.second.size());
          for (unsigned i : kv.second) put(out, (int32_t)i);
        }
        // ORBMatcher::searchByBow of the frame against itself as the "keyframe" (loop mode: no map-point masks)
        std::vector<cv::DMatch> bm;
        ORBMatcher bowMatcher(0.6f, true);
        int nbm = bowMatcher.searchByBow(f, voc, f->getLeftKeyPoints(), f->getLeftDescriptor(), fv, {}, {}, bm, false, true);
        put(out, (int32_t)nbm);
        for (auto &m : bm) put(out, (int32_t)m.queryIdx), put(out, (int32_t)m.trainIdx), put(out, m.distance);
      }
    }
    else if (mode == "threads")
    {
      // the reference's own pattern (src/Frame.cc:100-105): two ORBExtractors of ONE configuration, extract() on two
      // std::threads at the same time, several rounds; then getPyramid() of both (their contexts have moved on since)
      Camera::set(718.856f, 718.856f, 607.1928f, 185.2157f, 0.537166f);
      cv::Mat a = read_raw(argv[8], w, h, CV_8U), b = read_raw(argv[9], w, h, CV_8U);
      ORBExtractor exA(a, nf, nl, scale, tmpl, 20, 7), exB(b, nf, nl, scale, tmpl, 20, 7);
      std::vector<cv::KeyPoint> kA, kB;
      std::vector<cv::Mat> dA, dB;
      for (int round = 0; round < 6; ++round)
      {
        std::thread tA([&]() { exA.extract(kA, dA); });
        std::thread tB([&]() { exB.extract(kB, dB); });
        tA.join();
        tB.join();
      }
      dump(out, kA, dA);
      dump(out, kB, dB);
      // getPyramid() is valid without (before / long after) extract(): a third extractor that never extracted
      ORBExtractor exC(b, nf, nl, scale, tmpl, 20, 7);
      for (const ORBExtractor *e : {&exA, &exB, &exC})
      {
        const std::vector<cv::Mat> &pyr = e->getPyramid();
        put(out, (int32_t)pyr.size());
        const cv::Mat &top = pyr.back();
        put(out, (int32_t)top.cols);
        put(out, (int32_t)top.rows);
        for (int r = 0; r < top.rows; ++r) out.write((const char *)top.ptr<uchar>(r), top.cols);
      }
      // a Frame that is no longer resident must refuse device-side queries instead of answering with the newer frame's data
      Frame::SharedPtr f1 = Frame::createStereo(a, b, nf, tmpl, 20, 7, nullptr, nl, scale);
      int32_t g1 = (int32_t)f1->getGrids().size();
      Frame::SharedPtr f2 = Frame::createStereo(b, a, nf, tmpl, 20, 7, nullptr, nl, scale);
      int32_t refused = 0;
      try
      {
        f1->getGrids();
      }
      catch (const ORBSlam2Error &)
      {
        refused = 1;
      }
      put(out, g1);
      put(out, refused);
      put(out, (int32_t)f2->getGrids().size());
    }
    else if (mode == "rgbd")
    {
      Camera::set(520.908620f, 521.007327f, 325.141442f, 249.701764f, 0.0767889f, {0.231222f, -0.784899f, -0.003257f, -0.000105f, 0.917205f});
      cv::Mat g = read_raw(argv[8], w, h, CV_8U), dep = read_raw(argv[9], w, h, CV_16U);
      Frame::SharedPtr f = Frame::createRGBD(g, dep, nf, tmpl, 20, 7, nullptr, (float)atof(argv[11]), nl, scale);
      dump(out, f->getLeftKeyPoints(), f->getLeftDescriptor());
      put(out, (int32_t)f->getN());
      for (double v : f->getRightU()) put(out, v);
      for (double v : f->getDepth()) put(out, v);
    }
    return 0;
  }
  catch (const std::exception &e)
  {
    std::fprintf(stderr, "shim_main: %s\n", e.what());
    return 2;
  }
}
