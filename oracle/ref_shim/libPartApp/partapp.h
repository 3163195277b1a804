// Stand-in for libPartApp/partapp.h -- TEST INFRASTRUCTURE: PartApp reduced to the members the reference's
// libPictStruct/objectdetect.h and objectdetect_findrot.cpp touch.  loadScoreGrid (file input) aborts.
#pragma once
#include <QString>
#include <cstdlib>
#include <vector>
#include <libAdaBoost/AdaBoost.h>
#include <libMisc/misc.hpp>
#include <libAnnotation/annotationlist.h>
#include <libMultiArray/multi_array_def.h>
#include <libPartDetect/AbcDetectorParam.pb.h>
#include <libPartDetect/PartConfig.pb.h>
#include <libPartDetect/PartWindowParam.pb.h>
#include <libPartDetect/partdef.h>
#include <libPictStruct/objectdetect_aux.hpp>
#include "ExpParam.pb.h"

void bbox_from_pos(const ExpParam &exp_param, const PartWindowParam::PartParam &part_param, int scaleidx, int rotidx, int ix,
                   int iy, PartBBox &bbox);
void bbox_from_pos(const PartWindowParam::PartParam &part_param, double scale, double rot, int ix, int iy, PartBBox &bbox);

QString complete_relative_path(QString qsInputFile, QString qsReferenceFile);
namespace filesys {
inline bool check_dir(QString) { return true; }
inline bool create_dir(QString) { return true; }
inline bool check_file(QString) { return false; }
inline bool copy_file(QString, QString) { return false; }
inline void split_filename(QString, QString &, QString &) {}
}  // namespace filesys

class PartApp {
 public:
  ExpParam m_exp_param;
  PartConfig m_part_conf;
  PartWindowParam m_window_param;
  AbcDetectorParam m_abc_param;
  AnnotationList m_test_annolist, m_train_annolist;
  int m_rootpart_idx = -1;
  bool m_bExternalClassDir = false, m_bExternalSamplesDir = false;
  QString m_qsExpParam;
  void loadScoreGrid(std::vector<std::vector<FloatGrid2> > &, int, int, bool, bool, QString, QString) const { abort(); }
};
